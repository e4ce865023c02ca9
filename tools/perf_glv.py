"""A/B: BLS12-381 G1 MSM with and without the endomorphism split, 2^20 pairs resident on the device (stage timings)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kyber_b200 import workload as wl
from kyber_b200.capi import Engine
from oracle import bls12381 as o

n = 1 << 20
eng = Engine(0)
a = wl.prng_scalars("b2k/c2-a", 4096, o.R)
base = eng.bls12381_g1_mul_batch_affine(wl.scalars_to_bytes(a), wl.G1_BLS12381_AFFINE * 4096)
pts = torch.frombuffer(bytearray(base * (n // 4096)), dtype=torch.uint8).cuda()
sc = torch.frombuffer(bytearray(wl.scalars_to_bytes(wl.prng_scalars("b2k/c2-s", n, o.R))), dtype=torch.uint8).cuda()
out = torch.zeros(256, dtype=torch.uint8, device="cuda")
res = {}
import ctypes
for glv, m in ((1, 0), (0, 0), (1, 4), (1, 8), (1, 16), (1, 32), (1, 64)):
    eng.set_msm_glv(bool(glv))
    eng.lib.b2k_set_msm_chunk(eng.h, m)
    for _ in range(3):
        eng.call_dev("b2k_bls12381_g1_msm_dev", n, sc.data_ptr(), pts.data_ptr(), out.data_ptr()); eng.synchronize()
    acc = None
    for _ in range(10):
        eng.call_dev("b2k_bls12381_g1_msm_dev", n, sc.data_ptr(), pts.data_ptr(), out.data_ptr()); eng.synchronize()
        t = eng.last_timings()
        acc = t if acc is None else [x + y for x, y in zip(acc, t)]
    names = ["load", "digits_hist", "scan", "scatter", "accumulate", "reduce_chunks", "window_sum", "final", "pipeline", "fixup"]
    print("glv", glv, "m", m, dict(zip(names, [round(v / 10, 3) for v in acc])), bytes(out[:48].cpu().numpy()).hex()[:16])
