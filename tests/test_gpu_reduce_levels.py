"""GPU parity of the two bucket-reduction schemes (include/b2kyber.h: b2k_set_msm_reduce): one level (chunk + small scalar
multiplication) and two levels (running sums twice), forced and automatic, across window widths and curves -- all must give the
oracle's bytes."""
import random

import pytest

from kyber_b200 import workload as wl
from oracle import bls12381 as o

pytestmark = pytest.mark.gpu


def test_reduce_levels_bls12381_g1(engine):
    rng = random.Random(31)
    n = 5000
    a = wl.prng_scalars("b2k/red-a", n, o.R)
    pts = engine.bls12381_g1_mul_batch_affine(wl.scalars_to_bytes(a), wl.G1_BLS12381_AFFINE * n)
    for ks in (wl.prng_scalars("b2k/red-s", n, o.R), [rng.randrange(1 << 128) for _ in range(n)], [o.R - 1] * n):
        want = o.g1_compress(o.g1_mul(wl.dot_mod(ks, a, o.R)))
        sb = wl.scalars_to_bytes(ks)
        try:
            for c in (0, 8, 13, 16):
                engine.set_msm_window(c)
                for levels, m1, m2 in ((0, 0, 0), (1, 0, 0), (2, 0, 0), (2, 2, 8), (2, 8, 2), (2, 16, 16), (2, 1, 4), (2, 4, 1)):
                    engine.set_msm_reduce(levels, m1, m2)
                    assert engine.bls12381_g1_msm(sb, pts) == want, (c, levels, m1, m2)
                    if levels == 2 and c and (m1 or 8) * (m2 or 4) <= 64:
                        assert engine.last_msm_plan()["reduce_levels"] == 2
        finally:
            engine.set_msm_window(0)
            engine.set_msm_reduce(0, 0, 0)


def test_reduce_levels_other_curves(engine):
    """the same templates over Fp2 (BLS12-381 G2) and the 8-limb field (bn254 G1), two levels forced at c = 13"""
    from oracle import bn254 as o4
    rng = random.Random(37)
    n = 600
    ks = [rng.randrange(o.R) for _ in range(n)]
    a = [rng.randrange(1, o.R) for _ in range(n)]
    sb = wl.scalars_to_bytes(ks)
    p2 = engine.bls12381_g2_mul_batch_affine(wl.scalars_to_bytes(a), o.g2_to_affine_bytes(o.G2) * n)
    want2 = o.g2_compress(o.g2_mul(wl.dot_mod(ks, a, o.R)))
    k4 = [rng.randrange(o4.ORDER) for _ in range(n)]
    a4 = [rng.randrange(1, o4.ORDER) for _ in range(n)]
    g4 = o4.g1_marshal(o4.G1) if hasattr(o4, "G1") else o4.g1_marshal(o4.g1_mul(1))
    p4 = engine.bn254_g1_mul_batch(b"".join(x.to_bytes(32, "big") for x in a4), g4 * n)
    want4 = o4.g1_marshal(o4.g1_mul(sum(x * y for x, y in zip(k4, a4)) % o4.ORDER))
    try:
        engine.set_msm_window(13)
        for levels in (1, 2):
            engine.set_msm_reduce(levels, 0, 0)
            assert engine.bls12381_g2_msm(sb, p2) == want2, levels
            assert engine.bn254_g1_msm(b"".join(x.to_bytes(32, "big") for x in k4), p4) == want4, levels
    finally:
        engine.set_msm_window(0)
        engine.set_msm_reduce(0, 0, 0)
