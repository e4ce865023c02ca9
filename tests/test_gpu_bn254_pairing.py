"""GPU parity: bn254 pairing (GT bytes), ValidatePairing booleans and G2 Point.Mul / MSM vs the oracle that follows
pairing/bn254/{optate,twist,gfp2,gfp6,gfp12}.go line by line (oracle/bn254_pairing.py).  Property tests after
pairing/bn254/suite_test.go:253-268 (tripartite DH) and the bilinearity checks of the BLS12-381 suite."""
import random

import pytest

from oracle import bn254 as c, bn254_pairing as b

pytestmark = pytest.mark.gpu


def test_pair_gt_bytes_match_oracle(engine):
    rng = random.Random(71)
    pairs = [(c.G1, b.G2)] + [(c.g1_mul(rng.randrange(1, c.ORDER)), b.g2_mul(rng.randrange(1, c.ORDER))) for _ in range(5)]
    pairs += [(None, b.G2), (c.G1, None)]
    gt = engine.bn254_pair(b"".join(c.g1_marshal(p) for p, _ in pairs), b"".join(b.g2_marshal(q) for _, q in pairs))
    for i, (p, q) in enumerate(pairs):
        assert gt[384 * i:384 * (i + 1)] == b.gt_to_bytes(b.pairing(p, q)), i


def test_tripartite_dh_and_checks(engine):
    rng = random.Random(72)
    x, y, z = (rng.randrange(1, c.ORDER) for _ in range(3))
    # e(xG1, yG2)^z == e(yG1... ) expressed through byte-equal pairings of pre-multiplied points
    g1 = b"".join(c.g1_marshal(p) for p in (c.g1_mul(x * z % c.ORDER), c.g1_mul(x), c.g1_mul(z)))
    g2 = b"".join(b.g2_marshal(q) for q in (b.g2_mul(y), b.g2_mul(y * z % c.ORDER), b.g2_mul(x * y % c.ORDER)))
    gt = engine.bn254_pair(g1, g2)
    assert gt[:384] == gt[384:768] == gt[768:]
    a1 = c.g1_marshal(c.g1_mul(x)) * 2
    a2 = b.g2_marshal(b.g2_mul(y)) * 2
    b1 = c.g1_marshal(c.g1_mul(x * y % c.ORDER)) + c.g1_marshal(c.g1_mul((x * y + 1) % c.ORDER))
    b2 = b.g2_marshal(b.G2) * 2
    assert engine.bn254_pairing_check(a1, a2, b1, b2) == b"\x01\x00"


def test_bn254_g2_mul_and_msm(engine):
    rng = random.Random(73)
    n = 16
    ks = [0, 1, c.ORDER - 1] + [rng.randrange(c.ORDER) for _ in range(n - 3)]
    pts = [b.g2_mul(rng.randrange(1, c.ORDER)) for _ in range(n)]
    pts[4] = None
    pts[6], ks[6] = pts[5], ks[5]
    sb = b"".join(k.to_bytes(32, "big") for k in ks)
    pb = b"".join(b.g2_marshal(p) for p in pts)
    out = engine.bn254_g2_mul_batch(sb, pb)
    acc = None
    for i in range(n):
        r = b.g2_mul(ks[i], pts[i])
        assert out[128 * i:128 * i + 128] == b.g2_marshal(r), i
        acc = b.g2_add(acc, r)
    assert engine.bn254_g2_msm(sb, pb) == b.g2_marshal(acc)
