"""A/B: register cap of the inversion kernel of the affine rounds (b2k_set_msm_occupancy 4 = uncapped, 5 = 96 registers, leaving
a block slot per SM for another MSM's product kernel) -- per-MSM time with NC MSMs in flight, BLS12-381 G1, 2^20 pairs resident.
Every variant must give the oracle's bytes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from kyber_b200 import workload as wl
from kyber_b200.capi import Engine
from oracle import bls12381 as o

n = 1 << 20
eng = Engine(0)
a = wl.prng_scalars("b2k/c2-a", n, o.R)
s = wl.prng_scalars("b2k/c2", n, o.R)
base = eng.bls12381_g1_mul_batch_affine(wl.scalars_to_bytes(a), wl.G1_BLS12381_AFFINE * n)
want = o.g1_compress(o.g1_mul(wl.dot_mod(s, a, o.R))).hex()
pts = torch.frombuffer(bytearray(base), dtype=torch.uint8).cuda()
sc = torch.frombuffer(bytearray(wl.scalars_to_bytes(s)), dtype=torch.uint8).cuda()
torch.cuda.synchronize()


def throughput(minb, NC, K=40):
    engs = [Engine(0) for _ in range(NC)]
    streams = [torch.cuda.Stream() for _ in range(NC)]
    outs = [torch.zeros(256, dtype=torch.uint8, device="cuda") for _ in range(NC)]
    torch.cuda.synchronize()
    for e, st in zip(engs, streams):
        e.set_stream(st.cuda_stream)
        e._check(e.lib.b2k_set_msm_occupancy(e.h, minb))

    def go(k):
        engs[k % NC].call_dev("b2k_bls12381_g1_msm_dev", n, sc.data_ptr(), pts.data_ptr(), outs[k % NC].data_ptr())
    for k in range(3 * NC):
        go(k)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e0.record(streams[0])
    for k in range(K):
        go(k)
    ends = []
    for st in streams:
        e = torch.cuda.Event(enable_timing=True); e.record(st); ends.append(e)
    torch.cuda.synchronize()
    ms = max(e0.elapsed_time(e) for e in ends) / K
    ok = all(bytes(x[:48].cpu().numpy()).hex() == want for x in outs)
    print("inversion kernel blocks/SM bound", minb, "in flight", NC, "ms/MSM", round(ms, 3), "OK" if ok else "MISMATCH", flush=True)
    for e in engs:
        e.close()


for rep in range(2):
    for NC in (1, 4):
        for minb in (4, 5):
            throughput(minb, NC)
