"""GPU parity: bn254 G1 Point.Mul / MSM (the RecoverCommit curve, share/poly.go:449-476 over
pairing/bn254) vs the Python oracle oracle/bn254.py (restates pairing/bn254/curve.go, point.go)."""
import random

import pytest

from kyber_b200 import workload as wl
from oracle import bn254 as o

pytestmark = pytest.mark.gpu


def test_bn254_mul_batch_and_msm(engine):
    rng = random.Random(41)
    n = 40
    ks = [0, 1, o.ORDER - 1] + [rng.randrange(o.ORDER) for _ in range(n - 3)]
    pts = [o.g1_mul(rng.randrange(1, o.ORDER)) for _ in range(n)]
    pts[4] = None
    pts[6], ks[6] = pts[5], ks[5]
    pts[8], ks[8] = o.g1_neg(pts[7]), ks[7]
    sb = b"".join(k.to_bytes(32, "big") for k in ks)
    pb = b"".join(o.g1_marshal(p) for p in pts)
    out = engine.bn254_g1_mul_batch(sb, pb)
    acc = None
    for i in range(n):
        r = o.g1_mul(ks[i], pts[i])
        assert out[64 * i:64 * i + 64] == o.g1_marshal(r), i
        acc = o.g1_add(acc, r)
    for c in (0, 6, 8, 13):
        engine.set_msm_window(c)
        try:
            assert engine.bn254_g1_msm(sb, pb) == o.g1_marshal(acc), c
        finally:
            engine.set_msm_window(0)


def test_bn254_scalar_range(engine):
    from kyber_b200 import B2KError
    with pytest.raises(B2KError):
        engine.bn254_g1_msm(o.ORDER.to_bytes(32, "big"), o.g1_marshal(o.G1))


@pytest.mark.parametrize("t,n_extra", [(5, 3), (64, 0), (1024, 0)])
def test_recover_commit_matches_oracle(engine, t, n_extra):
    """share.RecoverCommit (share/poly.go:449-476) over bn254 G1: shares of a random degree-(t-1) polynomial
    recover f(0)*G (BASELINE configs[3] at t = 1024); indices arrive unsorted and with gaps."""
    from oracle import share_poly
    rng = random.Random(60 + t)
    coeffs = [rng.randrange(o.ORDER) for _ in range(t)]

    def f(x):
        acc = 0
        for c in reversed(coeffs):
            acc = (acc * x + c) % o.ORDER
        return acc
    idx = rng.sample(range(0, 3 * t), t + n_extra)
    # share values f(I+1)*G made by the engine itself (fixed-base batch), spot-checked against the oracle
    ys = [f(i + 1) for i in idx]
    gen = o.g1_marshal(o.G1)
    pts = engine.bn254_g1_mul_batch(b"".join(y.to_bytes(32, "big") for y in ys), gen * len(idx))
    assert pts[:64] == o.g1_marshal(o.g1_mul(ys[0]))
    shares = [(i, o.g1_unmarshal(pts[64 * k:64 * k + 64])) for k, i in enumerate(idx)]
    chosen = sorted(shares, key=lambda s: s[0])[:t]            # xyCommit: sort by index, first t
    got = engine.bn254_recover_commit([s[0] for s in chosen], b"".join(o.g1_marshal(s[1]) for s in chosen))
    assert got == o.g1_marshal(o.g1_mul(coeffs[0]))            # f(0)*G
    if t <= 64:
        assert got == o.g1_marshal(share_poly.recover_commit(o, shares, t))


def test_recover_commit_rejects_duplicate_index(engine):
    from kyber_b200 import B2KError
    g = o.g1_marshal(o.G1)
    with pytest.raises(B2KError):
        engine.bn254_recover_commit([1, 2, 2], g * 3)
