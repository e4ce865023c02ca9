"""A/B: bucket reduction of the BLS12-381 G1 MSM (2^20 pairs resident on the device) in one level (chunks of m buckets with a
small scalar multiplication per chunk) or two levels (m1, m2) -- stage timings from the library's CUDA events; every variant
must give the oracle's bytes.  Also times the throughput with 3 MSMs in flight for the default and the best variants."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
out = torch.zeros(256, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
names = ["load", "digits_hist", "scan", "scatter", "accumulate", "reduce_chunks", "window_sum", "final", "pipeline", "fixup", "rounds"]


def run(levels, m1, m2, K=6):
    eng.set_msm_reduce(levels, m1, m2)
    acc = None
    for k in range(K + 3):
        eng.call_dev("b2k_bls12381_g1_msm_dev", n, sc.data_ptr(), pts.data_ptr(), out.data_ptr()); eng.synchronize()
        if k >= 3:
            t = eng.last_timings()
            acc = t if acc is None else [x + y for x, y in zip(acc, t)]
    got = bytes(out[:48].cpu().numpy()).hex()
    d = dict(zip(names, [round(v / K, 3) for v in acc]))
    p = eng.last_msm_plan()
    print("levels", levels, "m1", m1, "m2", m2, "OK" if got == want else "MISMATCH", p["reduce_levels"], p["reduce_chunks"],
          {k: d[k] for k in ("reduce_chunks", "window_sum", "final", "pipeline")}, flush=True)


def throughput(levels, m1, m2, NC=3, K=30):
    engs = [Engine(0) for _ in range(NC)]
    streams = [torch.cuda.Stream() for _ in range(NC)]
    outs = [torch.zeros(256, dtype=torch.uint8, device="cuda") for _ in range(NC)]
    torch.cuda.synchronize()
    for e, st in zip(engs, streams):
        e.set_stream(st.cuda_stream)
        e.set_msm_reduce(levels, m1, m2)
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
    print("throughput levels", levels, "m1", m1, "m2", m2, "ms/MSM", round(ms, 3), "OK" if ok else "MISMATCH", flush=True)
    for e in engs:
        e.close()


run(1, 0, 0)
for m1, m2 in [(4, 4), (4, 8), (8, 4), (8, 8), (2, 8), (4, 2), (16, 4), (2, 4), (4, 16), (8, 16)]:
    run(2, m1, m2)
run(0, 0, 0)
throughput(1, 0, 0)
throughput(2, 4, 4)
throughput(2, 8, 4)
throughput(2, 4, 8)
throughput(2, 8, 8)
