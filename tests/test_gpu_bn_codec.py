"""GPU parity: batched UnmarshalBinary validation on the BN curves against the oracle's restatement of
the reference's acceptance rules (oracle/bn_codec.py: pairing/bn254/point.go:146-185, 473-514; pairing/bn256/point.go:206-238,
469-506): valid points, infinity, off-curve points, coordinates equal to / above the modulus (rejected on bn254, reduced
on bn256), and -- bn254 G2 only -- a point of the twist outside the order-n subgroup."""
import random

import pytest

from oracle import bn254 as o4
from oracle import bn254_pairing as p4
from oracle import bn256 as o6
from oracle import bn_codec as oc

pytestmark = pytest.mark.gpu


def _f2_sqrt(a, P, mul, pw):
    """square root in Fp[i]/(i^2+1), p = 3 mod 4 (None when a is not a square)"""
    if a == (0, 0):
        return a
    a1 = pw(a, (P - 3) // 4)
    alpha = mul(a1, mul(a1, a))
    x0 = mul(a1, a)
    if alpha == (P - 1, 0):
        r = ((-x0[1]) % P, x0[0])
    else:
        b = pw(((1 + alpha[0]) % P, alpha[1]), (P - 1) // 2)
        r = mul(b, x0)
    return r if mul(r, r) == (a[0] % P, a[1] % P) else None


def _pw(mul, P):
    def pw(a, e):
        r = (1, 0)
        for bit in bin(e)[2:]:
            r = mul(r, r)
            if bit == "1":
                r = mul(r, a)
        return r
    return pw


def _twist_point(rng, P, mul, add, twist_b):
    pw = _pw(mul, P)
    while True:
        x = (rng.randrange(P), rng.randrange(P))
        y = _f2_sqrt(add(mul(mul(x, x), x), twist_b), P, mul, pw)
        if y is not None:
            return (x, y)


def _g2_bytes(pt):
    (xr, xi), (yr, yi) = pt
    return b"".join(v.to_bytes(32, "big") for v in (xi, xr, yi, yr))


def test_bn254_unmarshal_checks(engine):
    rng = random.Random(91)
    P = o4.P
    g1 = [o4.g1_marshal(o4.g1_mul(rng.randrange(1, o4.ORDER))) for _ in range(6)]
    cases = g1 + [bytes(64)]
    cases.append(g1[0][:32] + (int.from_bytes(g1[0][32:], "big") ^ 1).to_bytes(32, "big"))        # off the curve
    cases.append(P.to_bytes(32, "big") + bytes(32))                                                 # x == p
    cases.append((int.from_bytes(g1[1][:32], "big") + P).to_bytes(32, "big") + g1[1][32:]
                 if int.from_bytes(g1[1][:32], "big") + P < 1 << 256 else g1[1])                    # x + p: same residue, rejected
    cases.append(bytes(32) + (P - 1).to_bytes(32, "big"))
    cases.append((1).to_bytes(32, "big") + (2).to_bytes(32, "big"))                                 # the generator (1, 2)
    got = engine.bn_unmarshal_check("bn254", 1, b"".join(cases))
    assert list(got) == [1 if oc.bn254_g1_ok(c) else 0 for c in cases]
    assert list(got[:7]) == [1] * 7 and got[7] == 0 and got[8] == 0 and got[-1] == 1

    g2 = [p4.g2_marshal(p4.g2_mul(rng.randrange(1, o4.ORDER))) for _ in range(3)]
    stray = _twist_point(rng, P, p4.f2_mul, p4.f2_add, p4.TWIST_B)                                  # on the twist, cofactor part alive
    cases2 = g2 + [bytes(128), _g2_bytes(stray)]
    cases2.append(g2[0][:96] + (int.from_bytes(g2[0][96:], "big") ^ 1).to_bytes(32, "big"))         # off the twist
    cases2.append(g2[1][:64] + P.to_bytes(32, "big") + g2[1][96:])                                  # a coordinate == p
    got2 = engine.bn_unmarshal_check("bn254", 2, b"".join(cases2))
    want2 = [1 if oc.bn254_g2_ok(c) else 0 for c in cases2]
    assert list(got2) == want2 and want2 == [1, 1, 1, 1, 0, 0, 0]


def test_bn256_unmarshal_checks(engine):
    rng = random.Random(92)
    P = o6.P
    g1 = [o6.g1_marshal(o6.g1_mul(rng.randrange(1, o6.ORDER))) for _ in range(5)]
    cases = g1 + [bytes(64), P.to_bytes(32, "big") + P.to_bytes(32, "big")]                         # (p, p) reads as infinity
    cases.append(g1[0][:32] + (int.from_bytes(g1[0][32:], "big") ^ 1).to_bytes(32, "big"))
    small = next(g for g in g1 if int.from_bytes(g[:32], "big") + P < 1 << 256)
    cases.append((int.from_bytes(small[:32], "big") + P).to_bytes(32, "big") + small[32:])          # x + p: accepted (no range check)
    got = engine.bn_unmarshal_check("bn256", 1, b"".join(cases))
    want = [1 if oc.bn256_g1_ok(c) else 0 for c in cases]
    assert list(got) == want and want == [1] * 7 + [0, 1]

    g2 = [o6.g2_marshal(o6.g2_mul(rng.randrange(1, o6.ORDER))) for _ in range(3)]
    stray = _twist_point(rng, P, o6.f2_mul, o6.f2_add, o6.TWIST_B)                                  # accepted: bn256 has no order check
    cases2 = g2 + [bytes(128), _g2_bytes(stray)]
    cases2.append(g2[0][:96] + (int.from_bytes(g2[0][96:], "big") ^ 1).to_bytes(32, "big"))
    got2 = engine.bn_unmarshal_check("bn256", 2, b"".join(cases2))
    want2 = [1 if oc.bn256_g2_ok(c) else 0 for c in cases2]
    assert list(got2) == want2 and want2 == [1, 1, 1, 1, 1, 0]
