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
for g in (1, 2, 4, 8):
    eng.set_msm_groups(g)
    for rep in range(3):
        t0 = time.perf_counter(); got = eng.bls12381_g1_msm(sb, pts); dt = time.perf_counter() - t0
    tm = eng.last_timings()
    print(f"groups={g} ok={got == exp} e2e_ms={dt*1e3:.2f} " + " ".join(f"{k}={v:.3f}" for k, v in zip(names, tm)), flush=True)
# skew + small sizes in grouped mode
import random
rng = random.Random(1)
for n2, dist in ((3000, "equal"), (3000, "bdn128"), (100000, "rand"), (17, "rand")):
    a2 = a[:n2]
    s2 = [0x1D2C3B4A59687] * n2 if dist == "equal" else ([rng.randrange(1 << 128) + 1 for _ in range(n2)] if dist == "bdn128" else s[:n2])
    p2 = pts[:96 * n2]
    want = o.g1_compress(o.g1_mul(wl.dot_mod(s2, a2, o.R)))
    for g in (1, 4):
        eng.set_msm_groups(g)
        assert eng.bls12381_g1_msm(wl.scalars_to_bytes(s2), p2) == want, (n2, dist, g)
print("grouped mode correct on skewed/small inputs")
