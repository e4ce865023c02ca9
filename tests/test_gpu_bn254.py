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
