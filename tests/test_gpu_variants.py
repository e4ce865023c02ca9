"""GPU parity of the tuning variants: every code layout / launch shape of the pairing kernels (b2k_set_pairing_variant: fused compact,
fused inlined, Miller + final exponentiation as two kernels) and every operand-staging mask of the affine rounds (b2k_set_msm_staging)
must give the bytes of the default path -- they are A/B knobs, not different algorithms.  Inputs include the exceptional cases the
reference's own tests use (infinity operands, wrong pairs) and, for the MSM, scalar sets that put equal points, P and -P and
infinities into one bucket (the group-law branches of the batched affine additions)."""
import random

import pytest

from oracle import bls12381 as o

pytestmark = pytest.mark.gpu

PAIRING_VARIANTS = (0, 1, 2, 4, 5, 6, 16, 17, 18, 19)


def test_every_pairing_variant_gives_the_same_bytes(engine):
    rng = random.Random(41)
    n = 70                                                   # more than one block of 64
    a1, a2, b1, b2, want = [], [], [], [], []
    for i in range(n):
        x, y = rng.randrange(1, o.R), rng.randrange(1, o.R)
        good = (i % 5 != 3)
        a1.append(o.g1_mul(x)); a2.append(o.g2_mul(y))
        b1.append(o.g1_mul(x * y % o.R if good else (x * y + 1) % o.R)); b2.append(o.G2)
        want.append(1 if good else 0)
    a1[7] = None; b1[7] = None                               # e(inf, Q) == e(inf, G2): both sides are 1
    want[7] = 1
    A1, A2 = b"".join(map(o.g1_to_affine_bytes, a1)), b"".join(map(o.g2_to_affine_bytes, a2))
    B1, B2 = b"".join(map(o.g1_to_affine_bytes, b1)), b"".join(map(o.g2_to_affine_bytes, b2))
    pairs = [(o.G1, o.G2), (None, o.G2), (o.G1, None)] + [(o.g1_mul(rng.randrange(1, o.R)), o.g2_mul(rng.randrange(1, o.R))) for _ in range(3)]
    g1 = b"".join(o.g1_to_affine_bytes(p) for p, _ in pairs)
    g2 = b"".join(o.g2_to_affine_bytes(q) for _, q in pairs)
    want_gt = b"".join(o.gt_to_bytes(o.pairing_reference(p, q)) for p, q in pairs)
    try:
        for v in PAIRING_VARIANTS:
            engine._check(engine.lib.b2k_set_pairing_variant(engine.h, v))
            assert list(engine.bls12381_pairing_check(A1, A2, B1, B2)) == want, v
            assert engine.bls12381_pair(g1, g2) == want_gt, v
        # a malformed operand (off the curve) makes its own check fail in every variant
        bad = bytearray(A1); bad[96 * 3 + 95] ^= 1
        for v in (0, 16, 19):
            engine._check(engine.lib.b2k_set_pairing_variant(engine.h, v))
            got = list(engine.bls12381_pairing_check(bytes(bad), A2, B1, B2))
            assert got[3] == 0 and got[:3] == want[:3] and got[4:] == want[4:], v
        assert engine.lib.b2k_set_pairing_variant(engine.h, 3) != 0 and engine.lib.b2k_set_pairing_variant(engine.h, 20) != 0
    finally:
        engine._check(engine.lib.b2k_set_pairing_variant(engine.h, 0))


def test_every_staging_mask_of_the_affine_rounds_gives_the_same_bytes(engine):
    rng = random.Random(42)
    n = 3000
    pts = [o.g1_mul(rng.randrange(1, o.R)) for _ in range(40)]
    pts[5] = None                                            # an operand at infinity
    pts[7] = pts[6]                                          # P twice
    pts[9] = o.g1_neg(pts[8])                                # P and -P
    plist = [pts[rng.randrange(40)] for _ in range(n)]
    pb = b"".join(o.g1_to_affine_bytes(p) for p in plist)
    sets = ([rng.randrange(o.R) for _ in range(n)],
            [0x1234567 + (i & 3) for i in range(n)])          # four scalar values: every bucket holds hundreds of (often equal) points
    try:
        engine.set_msm_window(8)
        engine.set_msm_affine(3, 16)                          # force three rounds of 16 outputs per thread on this small problem
        for ks in sets:
            sb = b"".join(o.scalar_to_bytes(k) for k in ks)
            want = o.g1_compress(o.g1_msm(ks, plist))
            for mask in range(16):
                engine._check(engine.lib.b2k_set_msm_staging(engine.h, mask))
                assert engine.bls12381_g1_msm(sb, pb) == want, mask
                plan = engine.last_msm_plan()
                assert plan["affine_rounds"] == 3 and plan["affine_split"], plan
    finally:
        engine._check(engine.lib.b2k_set_msm_staging(engine.h, 0))
        engine.set_msm_affine(-1, 0)
        engine.set_msm_window(0)
