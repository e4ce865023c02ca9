import sys, time
sys.path.insert(0, '.')
from kyber_b200 import Engine, workload as wl
from oracle import bls12381 as o
eng = Engine(0)
for logn in (16, 20):
    n = 1 << logn
    t = time.time()
    a = wl.prng_scalars("b2k/c2-a", n, o.R); s = wl.prng_scalars("b2k/c2", n, o.R)
    print("gen scalars", time.time() - t, flush=True)
    t = time.time()
    pts = eng.bls12381_g1_mul_batch_affine(wl.scalars_to_bytes(a), wl.G1_BLS12381_AFFINE * n)
    dt = time.time() - t
    print(f"mul_batch n={n}: {dt:.3f}s -> {n/dt:.3e} muls/s (e2e incl copies)", flush=True)
    sb = wl.scalars_to_bytes(s)
    for c in (0, 13, 14, 15, 16):
        eng.set_msm_window(c)
        for rep in range(2):
            t = time.time()
            got = eng.bls12381_g1_msm(sb, pts)
            dt = time.time() - t
        tm = eng.last_timings()
        print(f"msm n={n} c={c}: e2e {dt*1e3:.2f} ms; stages(ms)=", [round(x, 3) for x in tm], flush=True)
    exp = o.g1_compress(o.g1_mul(wl.dot_mod(s, a, o.R)))
    print("msm correct:", got == exp, flush=True)
