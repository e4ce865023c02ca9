"""GPU timing of the BLS12-381 G2 paths (MSM with 128-bit and full scalars, Point.Mul batch), checked against the oracle.
Usage (under gpurun): python tools/perf_g2.py > gpurun_out/<tag>_g2.txt"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kyber_b200 import Engine, workload as wl
from oracle import bls12381 as o

eng = Engine(0)
for logn in (16, 18):
    n = 1 << logn
    sks = wl.prng_scalars("b2k/g2perf", n, o.R)
    sb = wl.scalars_to_bytes(sks)
    pk_aff = eng.bls12381_g2_mul_batch_affine(sb, o.g2_to_affine_bytes(o.G2) * n)
    for name, ks in (("full scalars", wl.prng_scalars("b2k/g2perf-k", n, o.R)), ("128-bit scalars", [k >> 127 for k in wl.prng_scalars("b2k/g2perf-k", n, o.R)])):
        kb = wl.scalars_to_bytes(ks)
        got = eng.bls12381_g2_msm(kb, pk_aff)
        assert got == o.g2_compress(o.g2_mul(wl.dot_mod(ks, sks, o.R))), "G2 MSM differs from the oracle"
        t0 = time.perf_counter()
        for _ in range(5):
            eng.bls12381_g2_msm(kb, pk_aff)
        ms = (time.perf_counter() - t0) * 1e3 / 5
        tm = eng.last_timings()
        names = ["load", "digits_hist", "scan", "scatter", "accumulate", "reduce", "window_sum", "final", "pipeline", "fixup", "rounds"]
        print(f"G2 MSM 2^{logn} {name}: {ms:.2f} ms per call (host buffers) -> {n / ms * 1e3:.3e} muls/s; plan {eng.last_msm_plan()}; "
              f"device stages ms { {k: round(v, 3) for k, v in zip(names, tm)} }", flush=True)
    t0 = time.perf_counter()
    for _ in range(3):
        eng.bls12381_g2_mul_batch_affine(sb, pk_aff)
    ms = (time.perf_counter() - t0) * 1e3 / 3
    print(f"G2 Point.Mul batch 2^{logn}: {ms:.2f} ms per call (host buffers) -> {n / ms * 1e3:.3e} muls/s", flush=True)
