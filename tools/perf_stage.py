"""A/B on the GPU: backward pass of the affine rounds with cp.async-staged operands (b2k_set_msm_staging mask: bit 0 = round 0, bit 1 = later
rounds) against the plain loads, 2^20 and 2^21 pairs.  Usage (under gpurun): python tools/perf_stage.py > gpurun_out/<tag>_stage_ab.txt"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kyber_b200 import Engine, workload as wl
from oracle import bls12381 as o

dev = torch.device("cuda", 0)
n = 1 << 20
eng = Engine(0)
a = wl.prng_scalars("b2k/c2-a", n, o.R)
s = wl.prng_scalars("b2k/c2", n, o.R)
d_a = torch.frombuffer(bytearray(wl.scalars_to_bytes(a)), dtype=torch.uint8).to(dev)
d_gen = torch.frombuffer(bytearray(wl.G1_BLS12381_AFFINE), dtype=torch.uint8).to(dev).repeat(n)
d_pts = torch.empty(n * 96, dtype=torch.uint8, device=dev)
eng.call_dev("b2k_bls12381_g1_mul_batch_affine_dev", n, d_a.data_ptr(), d_gen.data_ptr(), d_pts.data_ptr())
d_s = torch.frombuffer(bytearray(wl.scalars_to_bytes(s)), dtype=torch.uint8).to(dev)
want = o.g1_compress(o.g1_mul(wl.dot_mod(s, a, o.R)))
NC = 4
streams = [torch.cuda.Stream(device=dev) for _ in range(NC)]
engs = [Engine(0) for _ in range(NC)]
outs = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(NC)]
for e, st in zip(engs, streams):
    e.set_stream(st.cuda_stream)
torch.cuda.synchronize()
for mask in (0, 15, 3, 12, 1, 4, 0, 15):
    for e in engs:
        e._check(e.lib.b2k_set_msm_staging(e.h, mask))
    for k in range(2 * NC):
        engs[k % NC].call_dev("b2k_bls12381_g1_msm_dev", n, d_s.data_ptr(), d_pts.data_ptr(), outs[k % NC].data_ptr())
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e0.record(streams[0])
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(NC)]
    K = 24
    for k in range(K):
        engs[k % NC].call_dev("b2k_bls12381_g1_msm_dev", n, d_s.data_ptr(), d_pts.data_ptr(), outs[k % NC].data_ptr())
    for st, ev in zip(streams, ends):
        ev.record(st)
    torch.cuda.synchronize()
    ms = max(e0.elapsed_time(ev) for ev in ends) / K
    for o_ in outs:
        assert bytes(o_[:48].cpu().tolist()) == want, "wrong MSM result with staging mask %d" % mask
    acc = []
    for _ in range(3):
        engs[0].call_dev("b2k_bls12381_g1_msm_dev", n, d_s.data_ptr(), d_pts.data_ptr(), outs[0].data_ptr())
        acc.append(engs[0].last_timings())
    tm = [sum(x[i] for x in acc) / 3 for i in range(len(acc[0]))]
    print(f"MSM 2^20 staging mask={mask}: {ms:.3f} ms per MSM pipelined ({n / ms * 1e3:.3e} muls/s); one MSM: accumulate {tm[4]:.3f} ms, rounds {tm[10]:.3f}, pipeline {tm[8]:.3f}", flush=True)
print("all masks give the oracle's result")
