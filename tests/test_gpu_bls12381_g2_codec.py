"""GPU parity: BLS12-381 G2 Point.Mul / MSM and batched UnmarshalBinary (decompress + subgroup check)
through the C ABI.  The decompression cases are the reference's own ZCash fixtures
(pairing/bls12381/deserialization_tests via tests/golden), exactly as TestZKCryptoVectorsG1Compressed /
G2Compressed drive them (pairing/bls12381/bls12381_test.go:74-186)."""
import json
import os
import random

import pytest

from kyber_b200 import workload as wl
from oracle import bls12381 as o

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
DESER = json.load(open(os.path.join(GOLD, "bls12381_deserialization.json")))


def _cases(group, size):
    out = []
    for v in DESER[group]:
        try:
            raw = bytes.fromhex(v["input"])
        except ValueError:
            continue
        if len(raw) == size:        # wrong-length vectors are rejected by the caller (fixed stride)
            out.append((v["name"], raw, v["valid"]))
    return out


def test_g1_decompress_zcash_fixtures(engine):
    cases = _cases("G1", 48)
    assert len(cases) >= 13
    out, ok = engine.bls12381_g1_decompress(b"".join(c[1] for c in cases))
    for i, (name, raw, valid) in enumerate(cases):
        assert bool(ok[i]) == valid, name
        if valid:
            assert out[96 * i:96 * i + 96] == o.g1_to_affine_bytes(o.g1_decompress(raw)), name


def test_g2_decompress_zcash_fixtures(engine):
    cases = _cases("G2", 96)
    assert len(cases) >= 15
    out, ok = engine.bls12381_g2_decompress(b"".join(c[1] for c in cases))
    for i, (name, raw, valid) in enumerate(cases):
        assert bool(ok[i]) == valid, name
        if valid:
            assert out[192 * i:192 * i + 192] == o.g2_to_affine_bytes(o.g2_decompress(raw)), name


def test_decompress_random_and_non_subgroup_points(engine):
    rng = random.Random(31)
    good1 = [o.g1_mul(rng.randrange(1, o.R)) for _ in range(20)]
    bad1 = []
    while len(bad1) < 12:                       # on the curve, outside the r-torsion subgroup
        x = rng.randrange(o.P)
        y = o.fp_sqrt((x ** 3 + 4) % o.P)
        if y is not None:
            bad1.append((x, y))
    data = b"".join(o.g1_compress(p) for p in good1 + bad1)
    out, ok = engine.bls12381_g1_decompress(data)
    assert list(ok) == [1] * 20 + [0] * 12
    for i, p in enumerate(good1):
        assert out[96 * i:96 * i + 96] == o.g1_to_affine_bytes(p)
    good2 = [o.g2_mul(rng.randrange(1, o.R)) for _ in range(10)]
    bad2 = []
    while len(bad2) < 6:
        x = (rng.randrange(o.P), rng.randrange(o.P))
        y = o.f2_sqrt(o.f2_add(o.f2_mul(o.f2_sqr(x), x), o.B2))
        if y is not None:
            bad2.append((x, y))
    out, ok = engine.bls12381_g2_decompress(b"".join(o.g2_compress(p) for p in good2 + bad2))
    assert list(ok) == [1] * 10 + [0] * 6
    for i, p in enumerate(good2):
        assert out[192 * i:192 * i + 192] == o.g2_to_affine_bytes(p)


def test_g2_mul_batch_and_msm(engine):
    rng = random.Random(32)
    n = 24
    ks = [0, 1, o.R - 1] + [rng.randrange(o.R) for _ in range(n - 3)]
    pts = [o.g2_mul(rng.randrange(1, o.R)) for _ in range(n)]
    pts[4] = None
    pts[6], ks[6] = pts[5], ks[5]
    pts[8], ks[8] = o.g2_neg(pts[7]), ks[7]
    sb = wl.scalars_to_bytes(ks)
    pb = b"".join(o.g2_to_affine_bytes(p) for p in pts)
    out = engine.bls12381_g2_mul_batch(sb, pb)
    acc = None
    for i in range(n):
        r = o.g2_mul(ks[i], pts[i])
        assert out[96 * i:96 * i + 96] == o.g2_compress(r), i
        acc = o.g2_add(acc, r)
    assert engine.bls12381_g2_msm(sb, pb) == o.g2_compress(acc)
    for c in (8, 13):
        engine.set_msm_window(c)
        try:
            assert engine.bls12381_g2_msm(sb, pb) == o.g2_compress(acc)
        finally:
            engine.set_msm_window(0)


def test_g2_msm_known_discrete_logs(engine):
    n = 2000
    a = wl.prng_scalars("b2k/test-g2a", n, o.R)
    s = wl.prng_scalars("b2k/test-g2s", n, o.R)
    pts = engine.bls12381_g2_mul_batch_affine(wl.scalars_to_bytes(a), o.g2_to_affine_bytes(o.G2) * n)
    assert pts[:192] == o.g2_to_affine_bytes(o.g2_mul(a[0]))
    assert engine.bls12381_g2_msm(wl.scalars_to_bytes(s), pts) == o.g2_compress(o.g2_mul(wl.dot_mod(s, a, o.R)))
