"""A/B: BLS12-381 G1 MSM, 2^20 distinct pairs resident on the device, with 0..R affine pair-tree rounds in front of the
XYZZ slices, fused or split rounds, several batch widths (stage timings from the library's CUDA events; every variant must
give the oracle's bytes).  `--profile`: run exactly one fused and one split single-round MSM (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kyber_b200 import workload as wl
from kyber_b200.capi import Engine
from oracle import bls12381 as o

profile = "--profile" in sys.argv
n = 1 << 20
eng = Engine(0)
a = wl.prng_scalars("b2k/c2-a", n, o.R)
s = wl.prng_scalars("b2k/c2", n, o.R)
base = eng.bls12381_g1_mul_batch_affine(wl.scalars_to_bytes(a), wl.G1_BLS12381_AFFINE * n)
want = o.g1_compress(o.g1_mul(wl.dot_mod(s, a, o.R))).hex()
pts = torch.frombuffer(bytearray(base), dtype=torch.uint8).cuda()
sc = torch.frombuffer(bytearray(wl.scalars_to_bytes(s)), dtype=torch.uint8).cuda()
out = torch.zeros(256, dtype=torch.uint8, device="cuda")
names = ["load", "digits_hist", "scan", "scatter", "accumulate", "reduce_chunks", "window_sum", "final", "pipeline", "fixup", "rounds"]


def run(split, rounds, batch, K):
    eng.set_msm_affine_split(bool(split))
    eng.set_msm_affine(rounds, batch)
    acc = None
    for k in range(K + (0 if profile else 3)):
        eng.call_dev("b2k_bls12381_g1_msm_dev", n, sc.data_ptr(), pts.data_ptr(), out.data_ptr()); eng.synchronize()
        if profile or k >= 3:
            t = eng.last_timings()
            acc = t if acc is None else [x + y for x, y in zip(acc, t)]
    got = bytes(out[:48].cpu().numpy()).hex()
    print("split", split, "rounds", rounds, "batch", batch, "OK" if got == want else "MISMATCH", eng.last_msm_plan()["affine_batch"],
          dict(zip(names, [round(v / K, 3) for v in acc])), flush=True)


if profile:
    run(0, 1, 64, 1)
    run(1, 1, 64, 1)
    sys.exit(0)
run(0, 0, 0, 8)
run(1, -1, 0, 8)
for split, rounds, batch in [(0, 2, 64), (1, 1, 64), (1, 2, 64), (1, 3, 64), (1, 1, 128), (1, 2, 128), (1, 3, 128), (1, 4, 128),
                             (1, 2, 256), (1, 3, 256), (1, 2, 0), (1, 3, 0), (1, 2, 512), (1, 5, 128), (1, 6, 128)]:
    run(split, rounds, batch, 8)
