"""GPU parity: BLS12-381 G1 Point.Mul batches and MSM through the C ABI vs the Python oracle.

Reads like the reference's own group tests (pairing/bls12381/bls12381_test.go:196-418 testGroup,
util/test/test.go:409-424 CompareGroups): same inputs through two implementations, compare
MarshalBinary bytes.
"""
import random

import pytest

from kyber_b200 import B2KError
from kyber_b200 import workload as wl
from oracle import bls12381 as o

pytestmark = pytest.mark.gpu


def _rand_points(rng, n):
    return [o.g1_mul(rng.randrange(1, o.R)) for _ in range(n)]


def test_mul_batch_matches_oracle(engine):
    rng = random.Random(11)
    n = 40
    ks = [0, 1, 2, o.R - 1, o.R - 2] + [rng.randrange(o.R) for _ in range(n - 5)]
    pts = _rand_points(rng, n)
    pts[7] = None                      # infinity operand
    out = engine.bls12381_g1_mul_batch(wl.scalars_to_bytes(ks), b"".join(o.g1_to_affine_bytes(p) for p in pts))
    for i in range(n):
        assert out[48 * i:48 * i + 48] == o.g1_compress(o.g1_mul(ks[i], pts[i])), i


def test_mul_batch_affine_output_roundtrip(engine):
    rng = random.Random(12)
    ks = [rng.randrange(o.R) for _ in range(9)] + [0]
    sb = wl.scalars_to_bytes(ks)
    out = engine.bls12381_g1_mul_batch_affine(sb, wl.G1_BLS12381_AFFINE * len(ks))
    for i, k in enumerate(ks):
        assert out[96 * i:96 * i + 96] == o.g1_to_affine_bytes(o.g1_mul(k)), i


def test_scalar_out_of_range_is_rejected(engine):
    bad = o.R.to_bytes(32, "big")     # == r: mod.Int.UnmarshalBinary rejects (group/mod/int.go:359-372)
    with pytest.raises(B2KError) as e:
        engine.bls12381_g1_mul_batch(bad, wl.G1_BLS12381_AFFINE)
    assert e.value.code == -3
    with pytest.raises(B2KError) as e:
        engine.bls12381_g1_msm(bad + (5).to_bytes(32, "big"), wl.G1_BLS12381_AFFINE * 2)
    assert e.value.code == -3


@pytest.mark.parametrize("c", [0, 4, 7, 8, 13, 16])
def test_msm_small_matches_oracle(engine, c):
    rng = random.Random(100 + c)
    n = 37
    ks = [rng.randrange(o.R) for _ in range(n)]
    pts = _rand_points(rng, n)
    # exceptional cases inside one bucket: P+P, P+(-P), infinity operand, zero scalar
    pts[1], ks[1] = pts[0], ks[0]
    pts[3], ks[3] = o.g1_neg(pts[2]), ks[2]
    pts[4] = None
    ks[5] = 0
    ks[6] = o.R - 1
    engine.set_msm_window(c)
    try:
        got = engine.bls12381_g1_msm(wl.scalars_to_bytes(ks), b"".join(o.g1_to_affine_bytes(p) for p in pts))
    finally:
        engine.set_msm_window(0)
    assert got == o.g1_compress(o.g1_msm(ks, pts))


def test_msm_single_and_all_cancel(engine):
    k = 0x1234567890ABCDEF
    p = o.g1_mul(77)
    assert engine.bls12381_g1_msm(wl.scalars_to_bytes([k]), o.g1_to_affine_bytes(p)) == o.g1_compress(o.g1_mul(k, p))
    # k*P + (r-k)*P = infinity -> c0 00 .. 00
    got = engine.bls12381_g1_msm(wl.scalars_to_bytes([k, o.R - k]), o.g1_to_affine_bytes(p) * 2)
    assert got == o.g1_compress(None)


@pytest.mark.parametrize("n", [1000, 1 << 14])
def test_msm_known_discrete_logs(engine, n):
    """sum s_i*(a_i*G) == ((sum s_i a_i) mod r)*G  (SURVEY 8c trick (i)); points made by the engine itself,
    a sample of them re-checked against the oracle."""
    a = wl.prng_scalars("b2k/test-a", n, o.R)
    s = wl.prng_scalars("b2k/test-s", n, o.R)
    pts = engine.bls12381_g1_mul_batch_affine(wl.scalars_to_bytes(a), wl.G1_BLS12381_AFFINE * n)
    for i in (0, 1, n // 2, n - 1):
        assert pts[96 * i:96 * i + 96] == o.g1_to_affine_bytes(o.g1_mul(a[i]))
    got = engine.bls12381_g1_msm(wl.scalars_to_bytes(s), pts)
    assert got == o.g1_compress(o.g1_mul(wl.dot_mod(s, a, o.R)))


@pytest.mark.parametrize("dist_name", ["equal", "small", "two_values", "bdn128"])
def test_msm_skewed_scalar_distributions(engine, dist_name):
    """Balanced-slice accumulate must not depend on the scalar distribution (buckets with thousands of
    points, empty upper windows, 128-bit BDN coefficients sign/bdn/bdn.go:29-63)."""
    n = 3000
    rng = random.Random(77)
    a = wl.prng_scalars("b2k/test-a", n, o.R)
    if dist_name == "equal":
        s = [0x1D2C3B4A59687] * n
    elif dist_name == "small":
        s = [rng.randrange(1 << 20) for _ in range(n)]
    elif dist_name == "two_values":
        s = [rng.choice([1, o.R - 1]) for _ in range(n)]
    else:
        s = [rng.randrange(1 << 128) + 1 for _ in range(n)]
    pts = engine.bls12381_g1_mul_batch_affine(wl.scalars_to_bytes(a), wl.G1_BLS12381_AFFINE * n)
    want = o.g1_compress(o.g1_mul(wl.dot_mod(s, a, o.R)))
    for c, L in ((0, 0), (16, 0), (8, 3), (11, 1)):
        engine.set_msm_window(c)
        engine.set_msm_slice(L)
        try:
            got = engine.bls12381_g1_msm(wl.scalars_to_bytes(s), pts)
        finally:
            engine.set_msm_window(0)
            engine.set_msm_slice(0)
        assert got == want, (dist_name, c, L)


@pytest.mark.parametrize("dist_name", ["uniform", "equal", "small", "two_values", "bdn128"])
def test_msm_affine_pair_tree_rounds_agree(engine, dist_name):
    """The affine pair-tree rounds (batched affine additions around one inversion per thread, msm_affine.cuh) in front
    of the XYZZ slices: forced on at small sizes, every round count and batch width, on uniform and skewed scalar sets
    with repeated points, P / -P pairs and operands at infinity in the same bucket -- same bytes as the oracle."""
    n = 3000
    rng = random.Random(79)
    a = wl.prng_scalars("b2k/test-a", n, o.R)
    a[10] = a[11] = a[12]                     # equal points (P + P inside a bucket when their scalars agree)
    a[20] = o.R - a[21]                       # P and -P
    a[30] = 0                                 # operand at infinity
    if dist_name == "uniform":
        s = wl.prng_scalars("b2k/test-aff", n, o.R)
        s[10] = s[11] = s[12]
        s[20] = s[21]
    elif dist_name == "equal":
        s = [0x1D2C3B4A59687] * n
    elif dist_name == "small":
        s = [rng.randrange(1 << 20) for _ in range(n)]
    elif dist_name == "two_values":
        s = [rng.choice([1, o.R - 1]) for _ in range(n)]
    else:
        s = [rng.randrange(1 << 128) + 1 for _ in range(n)]
    pts = engine.bls12381_g1_mul_batch_affine(wl.scalars_to_bytes(a), wl.G1_BLS12381_AFFINE * n)
    want = o.g1_compress(o.g1_mul(wl.dot_mod(s, a, o.R)))
    sb = wl.scalars_to_bytes(s)
    for split in (True, False):               # every round as three kernels (default) / as one fused kernel
        for rounds, batch, c, L in ((1, 8, 0, 0), (2, 33, 8, 3), (3, 64, 6, 0), (5, 1, 5, 2), (8, 17, 4, 0), (4, 0, 16, 0), (2, 300, 7, 0)):
            engine.set_msm_affine_split(split)
            engine.set_msm_affine(rounds, batch)
            engine.set_msm_window(c)
            engine.set_msm_slice(L)
            try:
                got = engine.bls12381_g1_msm(sb, pts)
            finally:
                engine.set_msm_affine_split(True)
                engine.set_msm_affine(-1, 0)
                engine.set_msm_window(0)
                engine.set_msm_slice(0)
            assert got == want, (dist_name, split, rounds, batch, c, L)
    engine.set_msm_affine(0, 0)               # off: the plain XYZZ pipeline
    try:
        assert engine.bls12381_g1_msm(sb, pts) == want
    finally:
        engine.set_msm_affine(-1, 0)


def test_msm_variants_agree(engine):
    n = 2000
    a = wl.prng_scalars("b2k/test-a", n, o.R)
    s = wl.prng_scalars("b2k/test-v", n, o.R)
    pts = engine.bls12381_g1_mul_batch_affine(wl.scalars_to_bytes(a), wl.G1_BLS12381_AFFINE * n)
    sb = wl.scalars_to_bytes(s)
    engine.set_msm_variant(True)
    try:
        v1 = engine.bls12381_g1_msm(sb, pts)
    finally:
        engine.set_msm_variant(False)
    assert v1 == engine.bls12381_g1_msm(sb, pts) == o.g1_compress(o.g1_mul(wl.dot_mod(s, a, o.R)))


@pytest.mark.parametrize("dist_name", ["uniform", "edges", "bdn128", "equal"])
def test_msm_endomorphism_split_on_and_off_agree(engine, dist_name):
    """The GLV front end (default) and the plain 255-bit pipeline give the same bytes as the oracle: uniform scalars,
    scalars at the split's edges (multiples of x^2, halves, r-1 ...), 128-bit BDN factors, one repeated scalar."""
    n = 4100
    rng = random.Random(78)
    x2 = o.X_ABS ** 2
    a = wl.prng_scalars("b2k/test-a", n, o.R)
    if dist_name == "uniform":
        s = wl.prng_scalars("b2k/test-g", n, o.R)
    elif dist_name == "edges":
        base = [0, 1, x2 - 1, x2, x2 + 1, x2 // 2, x2 // 2 + 1, (o.R - 1) // 2, (o.R + 1) // 2, o.R - 1, o.R - x2, o.R - x2 // 2, (1 << 254) - 3]
        s = [base[i % len(base)] + (i // len(base)) * x2 for i in range(n)]
        s = [v % o.R for v in s]
    elif dist_name == "bdn128":
        s = [rng.randrange(1 << 128) + 1 for _ in range(n)]
    else:
        s = [0x5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A5A] * n
    pts = engine.bls12381_g1_mul_batch_affine(wl.scalars_to_bytes(a), wl.G1_BLS12381_AFFINE * n)
    want = o.g1_compress(o.g1_mul(wl.dot_mod(s, a, o.R)))
    sb = wl.scalars_to_bytes(s)
    for glv in (True, False):
        for c in (0, 7, 16):
            engine.set_msm_glv(glv)
            engine.set_msm_window(c)
            try:
                got = engine.bls12381_g1_msm(sb, pts)
            finally:
                engine.set_msm_glv(True)
                engine.set_msm_window(0)
            assert got == want, (dist_name, glv, c)


def test_msm_async_two_contexts_and_deferred_status(engine):
    """b2k_bls12381_g1_msm_async + b2k_wait: two contexts driven alternately by one thread give the oracle's bytes for
    every submitted batch; an out-of-range scalar surfaces at b2k_wait (as B2K_ERR_SCALAR_RANGE), not at submission."""
    import ctypes
    from kyber_b200.capi import Engine, B2KError
    other = Engine(0)
    engs = [engine, other]
    n = 1500
    a = wl.prng_scalars("b2k/test-a", n, o.R)
    pts = engine.bls12381_g1_mul_batch_affine(wl.scalars_to_bytes(a), wl.G1_BLS12381_AFFINE * n)
    batches, outs, wants = [], [], []
    for k in range(5):
        s = wl.prng_scalars("b2k/test-async-%d" % k, n, o.R)
        batches.append(ctypes.create_string_buffer(wl.scalars_to_bytes(s), 32 * n))
        outs.append(ctypes.create_string_buffer(48))
        wants.append(o.g1_compress(o.g1_mul(wl.dot_mod(s, a, o.R))))
    pbuf = ctypes.create_string_buffer(pts, 96 * n)
    for k in range(5):
        e = engs[k % 2]
        if k >= 2:
            e._check(e.lib.b2k_wait(e.h))
            assert outs[k - 2].raw == wants[k - 2]
        e._check(e.lib.b2k_bls12381_g1_msm_async(e.h, n, batches[k], pbuf, outs[k]))
    for k in (3, 4):
        e = engs[k % 2]
        e._check(e.lib.b2k_wait(e.h))
        assert outs[k].raw == wants[k]
    bad = ctypes.create_string_buffer(o.R.to_bytes(32, "big") + bytes(32 * (n - 1)), 32 * n)
    engine._check(engine.lib.b2k_bls12381_g1_msm_async(engine.h, n, bad, pbuf, outs[0]))      # accepted ...
    with pytest.raises(B2KError) as ei:
        engine._check(engine.lib.b2k_wait(engine.h))                                           # ... reported here
    assert ei.value.code == -3
