"""GPU parity for the target group as a kyber.Group, the Miller / Finalize split, products of pairings and batched Point.Add
(include/b2kyber.h: b2k_*_gt_{mul,inv,exp}, b2k_*_miller, b2k_*_final_exp / finalize, b2k_*_pairing_product_check,
b2k_bls12381_g{1,2}_add_batch) against the oracles.
Reference semantics: pairing/bls12381/kilic/gt.go:33-83 (Add = product, Neg = inverse, Mul = exponentiation),
pairing/bn254/point.go:560-623, 768-786 (pointGT.Add/Neg/Mul, Miller, Finalize), SURVEY.md 8e (one final exponentiation)."""
import random

import pytest

from oracle import bls12381 as o
from oracle import bn254 as c4, bn254_pairing as b4

pytestmark = pytest.mark.gpu


def _bls_pairs(rng, n):
    return [(o.g1_mul(rng.randrange(1, o.R)), o.g2_mul(rng.randrange(1, o.R))) for _ in range(n)]


def test_bls12381_gt_group_operations_match_the_oracle(engine):
    rng = random.Random(201)
    pairs = _bls_pairs(rng, 4)
    g1 = b"".join(o.g1_to_affine_bytes(p) for p, _ in pairs)
    g2 = b"".join(o.g2_to_affine_bytes(q) for _, q in pairs)
    gt = engine.bls12381_pair(g1, g2)
    vals = [o.pairing_reference(p, q) for p, q in pairs]
    assert gt == b"".join(o.gt_to_bytes(v) for v in vals)
    # Add = Fp12 product (gt.go:59-64), Null = 1
    one = o.gt_to_bytes(o.F12_ONE)
    a, b = gt, gt[576:] + one
    want = [o.f12_mul(vals[i], vals[i + 1]) for i in range(3)] + [vals[3]]
    assert engine.gt_mul("bls12381", a, b) == b"".join(o.gt_to_bytes(v) for v in want)
    # Neg = inverse (gt.go:71-75): a * a^-1 = 1
    inv = engine.gt_inv("bls12381", gt)
    assert inv == b"".join(o.gt_to_bytes(o.f12_inv(v)) for v in vals)
    assert engine.gt_mul("bls12381", gt, inv) == one * 4
    # Mul = exponentiation (gt.go:77-83)
    ks = [0, 1, o.R - 1, rng.randrange(o.R)]
    got = engine.gt_exp("bls12381", b"".join(o.scalar_to_bytes(k) for k in ks), gt)
    assert got == b"".join(o.gt_to_bytes(o.f12_pow(v, k)) for v, k in zip(vals, ks))
    # bilinearity through the group interface: e(P, Q)^k == e(kP, Q)
    k = ks[3]
    assert got[3 * 576:] == engine.bls12381_pair(o.g1_to_affine_bytes(o.g1_mul(k, pairs[3][0])), o.g2_to_affine_bytes(pairs[3][1]))


def test_bls12381_gt_rejects_non_canonical_coefficients(engine):
    from kyber_b200 import B2KError
    bad = (o.P).to_bytes(48, "big") + bytes(576 - 48)                 # first coefficient == p
    with pytest.raises(B2KError) as ei:
        engine.gt_inv("bls12381", bad)
    assert ei.value.code == -5


def test_bls12381_miller_then_final_exp_is_pair(engine):
    rng = random.Random(202)
    pairs = _bls_pairs(rng, 3) + [(None, o.G2), (o.G1, None)]
    g1 = b"".join(o.g1_to_affine_bytes(p) for p, _ in pairs)
    g2 = b"".join(o.g2_to_affine_bytes(q) for _, q in pairs)
    f = engine.miller("bls12381", g1, g2)
    e = engine.final_exp("bls12381", f)
    assert e == engine.bls12381_pair(g1, g2)
    assert e[3 * 576:] == o.gt_to_bytes(o.F12_ONE) * 2             # infinity operand -> 1
    # several Miller values multiplied, ONE final exponentiation == product of the pairings
    prod = f[:576]
    for i in range(1, 3):
        prod = engine.gt_mul("bls12381", prod, f[576 * i:576 * (i + 1)])
    want = o.F12_ONE
    for p, q in pairs[:3]:
        want = o.f12_mul(want, o.pairing_reference(p, q))
    assert engine.final_exp("bls12381", prod) == o.gt_to_bytes(want)
    assert engine.bls12381_pairing_product(g1[:3 * 96], g2[:3 * 192]) == o.gt_to_bytes(want)


@pytest.mark.parametrize("n", [1, 2, 3, 33, 130])
def test_bls12381_pairing_product_check(engine, n):
    """prod e(a_i G1, b_i G2) * e(-(sum a_i b_i) G1, G2) == 1; a perturbed member breaks it; n spans several blocks"""
    rng = random.Random(203 + n)
    ab = [(rng.randrange(1, o.R), rng.randrange(1, o.R)) for _ in range(n - 1)]
    tot = sum(a * b for a, b in ab) % o.R
    scal_a = b"".join(o.scalar_to_bytes(a) for a, _ in ab) + o.scalar_to_bytes((o.R - tot) % o.R)
    g1 = engine.bls12381_g1_mul_batch_affine(scal_a, o.g1_to_affine_bytes(o.G1) * n)
    scal_b = b"".join(o.scalar_to_bytes(b) for _, b in ab) + o.scalar_to_bytes(1)
    g2 = engine.bls12381_g2_mul_batch_affine(scal_b, o.g2_to_affine_bytes(o.G2) * n)
    if n == 1:                                                      # e(0 G1, G2) = e(inf, G2) = 1
        assert g1 == bytes(96)
    assert engine.pairing_product_check("bls12381", g1, g2) is True
    if n > 1:
        bad = bytearray(g1)
        bad[:96] = o.g1_to_affine_bytes(o.g1_mul(ab[0][0] + 1))
        assert engine.pairing_product_check("bls12381", bytes(bad), g2) is False


def test_bn254_miller_finalize_product_and_gt_ops(engine):
    rng = random.Random(204)
    ks = [(rng.randrange(1, c4.ORDER), rng.randrange(1, c4.ORDER)) for _ in range(3)]
    pairs = [(c4.g1_mul(a), b4.g2_mul(b)) for a, b in ks] + [(None, b4.G2)]
    g1 = b"".join(c4.g1_marshal(p) for p, _ in pairs)
    g2 = b"".join(b4.g2_marshal(q) for _, q in pairs)
    f = engine.miller("bn254", g1, g2)
    e = engine.final_exp("bn254", f)                                 # pointGT.Finalize(pointGT.Miller(...)) == Pair(...)
    assert e == engine.bn254_pair(g1, g2)
    vals = [b4.pairing(p, q) for p, q in pairs]
    assert e == b"".join(b4.gt_to_bytes(v) for v in vals)
    assert engine.gt_mul("bn254", e[:384 * 3], e[384:]) == b"".join(b4.gt_to_bytes(b4.f12_mul(vals[i], vals[i + 1])) for i in range(3))
    assert engine.gt_inv("bn254", e) == b"".join(b4.gt_to_bytes(b4.f12_inv(v)) for v in vals)
    xs = [0, 1, c4.ORDER - 1, rng.randrange(c4.ORDER)]
    assert engine.gt_exp("bn254", b"".join(x.to_bytes(32, "big") for x in xs), e) == \
        b"".join(b4.gt_to_bytes(b4.f12_pow(v, x)) for v, x in zip(vals, xs))
    # product check with one final exponentiation
    tot = sum(a * b for a, b in ks) % c4.ORDER
    g1c = g1[:3 * 64] + c4.g1_marshal(c4.g1_mul((c4.ORDER - tot) % c4.ORDER))
    g2c = g2[:3 * 128] + b4.g2_marshal(b4.G2)
    assert engine.pairing_product_check("bn254", g1c, g2c) is True
    assert engine.pairing_product_check("bn254", g1c, g2[:3 * 128] + b4.g2_marshal(b4.g2_mul(2))) is False


def test_bn256_miller_finalize_and_product(engine):
    from oracle import bn256 as c6
    rng = random.Random(205)
    a, b = rng.randrange(1, c6.ORDER), rng.randrange(1, c6.ORDER)
    g1 = c6.g1_marshal(c6.g1_mul(a)) + c6.g1_marshal(c6.g1_mul((c6.ORDER - a * b) % c6.ORDER))
    g2 = c6.g2_marshal(c6.g2_mul(b)) + c6.g2_marshal(c6.G2)
    assert engine.final_exp("bn256", engine.miller("bn256", g1, g2)) == engine.bn256_pair(g1, g2)
    assert engine.pairing_product_check("bn256", g1, g2) is True
    assert engine.pairing_product_check("bn256", g1[:64] * 2, g2) is False


def test_bls12381_point_add_batches(engine):
    rng = random.Random(206)
    P = [o.g1_mul(rng.randrange(1, o.R)) for _ in range(4)]
    A = [P[0], P[1], P[2], None, P[3], P[3]]
    B = [P[1], P[1], None, P[2], o.g1_neg(P[3]), None]
    a = b"".join(o.g1_to_affine_bytes(p) for p in A)
    b = b"".join(o.g1_to_affine_bytes(p) for p in B)
    assert engine.bls12381_add_batch(1, a, b) == b"".join(o.g1_to_affine_bytes(o.g1_add(x, y)) for x, y in zip(A, B))
    assert engine.bls12381_add_batch(1, a, b, negate_b=True) == b"".join(o.g1_to_affine_bytes(o.g1_add(x, o.g1_neg(y))) for x, y in zip(A, B))
    Q = [o.g2_mul(rng.randrange(1, o.R)) for _ in range(3)]
    A2 = [Q[0], Q[1], None, Q[2]]
    B2 = [Q[1], Q[1], Q[2], o.g2_neg(Q[2])]
    a2 = b"".join(o.g2_to_affine_bytes(p) for p in A2)
    b2 = b"".join(o.g2_to_affine_bytes(p) for p in B2)
    assert engine.bls12381_add_batch(2, a2, b2) == b"".join(o.g2_to_affine_bytes(o.g2_add(x, y)) for x, y in zip(A2, B2))


def test_operands_off_the_curve_or_out_of_range_are_refused(engine):
    """B2K_ERR_POINT is a promise of the header -- an off-curve or non-canonical operand must not be multiplied"""
    from kyber_b200 import B2KError
    good = o.g1_to_affine_bytes(o.G1)
    off = good[:95] + bytes([good[95] ^ 1])                          # y perturbed: not on the curve
    big = (o.P + o.G1_X).to_bytes(48, "big") + good[48:]             # x + p: same residue, non-canonical bytes
    one = o.scalar_to_bytes(5)
    for badpt in (off, big):
        for call in (lambda: engine.bls12381_g1_mul_batch(one, badpt),
                     lambda: engine.bls12381_g1_msm(one * 3, good + badpt + good),
                     lambda: engine.bls12381_pair(badpt, o.g2_to_affine_bytes(o.G2))):
            with pytest.raises(B2KError) as ei:
                call()
            assert ei.value.code == -5
    # the status word does not stick: a clean call after the failure succeeds
    assert engine.bls12381_g1_mul_batch(one, good) == o.g1_compress(o.g1_mul(5))
    # a pairing CHECK reports the malformed element as a failed check, like ValidatePairing's boolean
    ok = engine.bls12381_pairing_check(good + off, o.g2_to_affine_bytes(o.G2) * 2, good * 2, o.g2_to_affine_bytes(o.G2) * 2)
    assert ok == b"\x01\x00"
    # bn254: coordinate >= p refused (gfP.Unmarshal), off-curve refused
    g = c4.g1_marshal(c4.G1)
    with pytest.raises(B2KError):
        engine.bn254_g1_mul_batch((3).to_bytes(32, "big"), (c4.P + 1).to_bytes(32, "big") + g[32:])
    with pytest.raises(B2KError):
        engine.bn254_g1_mul_batch((3).to_bytes(32, "big"), g[:63] + bytes([g[63] ^ 1]))
