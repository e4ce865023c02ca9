import sys, time
sys.path.insert(0, '.')
from kyber_b200 import Engine, workload as wl
from oracle import bls12381 as o
eng = Engine(0)
n = 1 << 20
a = wl.prng_scalars("b2k/c2-a", n, o.R); s = wl.prng_scalars("b2k/c2", n, o.R)
pts = eng.bls12381_g1_mul_batch_affine(wl.scalars_to_bytes(a), wl.G1_BLS12381_AFFINE * n)
sb = wl.scalars_to_bytes(s)
exp = o.g1_compress(o.g1_mul(wl.dot_mod(s, a, o.R)))
names = ["load", "count", "scan", "scatter", "accum", "reduce", "wsum", "final", "total", "fixup/tail"]
for m in (0, 4, 8, 16, 32, 64, 128):
    eng._check(eng.lib.b2k_set_msm_chunk(eng.h, m))
    for rep in range(3):
        got = eng.bls12381_g1_msm(sb, pts)
    tm = eng.last_timings()
    print(f"m={m} ok={got == exp} " + " ".join(f"{k}={v:.3f}" for k, v in zip(names, tm)), flush=True)
