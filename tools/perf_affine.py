"""A/B: BLS12-381 G1 MSM, 2^20 distinct pairs resident on the device, with 0..R affine pair-tree rounds in front of the
XYZZ slices and several batch widths (stage timings from the library's CUDA events; every variant must give the oracle's bytes)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kyber_b200 import workload as wl
from kyber_b200.capi import Engine
from oracle import bls12381 as o

n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
eng = Engine(0)
a = wl.prng_scalars("b2k/c2-a", n, o.R)
s = wl.prng_scalars("b2k/c2", n, o.R)
base = eng.bls12381_g1_mul_batch_affine(wl.scalars_to_bytes(a), wl.G1_BLS12381_AFFINE * n)
want = o.g1_compress(o.g1_mul(wl.dot_mod(s, a, o.R))).hex()
pts = torch.frombuffer(bytearray(base), dtype=torch.uint8).cuda()
sc = torch.frombuffer(bytearray(wl.scalars_to_bytes(s)), dtype=torch.uint8).cuda()
out = torch.zeros(256, dtype=torch.uint8, device="cuda")
names = ["load", "digits_hist", "scan", "scatter", "accumulate", "reduce_chunks", "window_sum", "final", "pipeline", "fixup", "rounds"]
cfgs = [(0, 0), (-1, 0)] + [(r, b) for r in (1, 2, 3, 4, 5) for b in (32, 48, 64)] + [(3, 16), (3, 24), (4, 24)]
for rounds, batch in cfgs:
    eng.set_msm_affine(rounds, batch)
    for _ in range(3):
        eng.call_dev("b2k_bls12381_g1_msm_dev", n, sc.data_ptr(), pts.data_ptr(), out.data_ptr()); eng.synchronize()
    acc = None
    K = 8
    for _ in range(K):
        eng.call_dev("b2k_bls12381_g1_msm_dev", n, sc.data_ptr(), pts.data_ptr(), out.data_ptr()); eng.synchronize()
        t = eng.last_timings()
        acc = t if acc is None else [x + y for x, y in zip(acc, t)]
    got = bytes(out[:48].cpu().numpy()).hex()
    print("rounds", rounds, "batch", batch, "OK" if got == want else "MISMATCH", dict(zip(names, [round(v / K, 3) for v in acc])), flush=True)
