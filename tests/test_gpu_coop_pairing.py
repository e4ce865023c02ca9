"""GPU parity of the warp-cooperative BLS12-381 pairing (coop_pairing.cuh; small batches of Suite.Pair / ValidatePairing,
kilic/suite.go:57-75): same GT bytes and booleans as the oracle and as the one-per-thread kernels, including infinity operands (a pair
with an infinity member contributes 1), wrong pairs, a malformed operand, shared second operands (the bls.Verify wrappers) and both
sides of the dispatch threshold."""
import random

import pytest

from oracle import bls12381 as o

pytestmark = pytest.mark.gpu


def _check_inputs(rng, n):
    a1, a2, b1, b2, want = [], [], [], [], []
    for i in range(n):
        x, y = rng.randrange(1, o.R), rng.randrange(1, o.R)
        good = (i % 5 != 3)
        a1.append(o.g1_mul(x)); a2.append(o.g2_mul(y))
        b1.append(o.g1_mul(x * y % o.R if good else (x * y + 1) % o.R)); b2.append(o.G2)
        want.append(1 if good else 0)
    return a1, a2, b1, b2, want


def test_coop_pair_and_check_match_the_oracle_and_the_batch_kernels(engine):
    rng = random.Random(61)
    pairs = [(o.G1, o.G2), (None, o.G2), (o.G1, None)] + [(o.g1_mul(rng.randrange(1, o.R)), o.g2_mul(rng.randrange(1, o.R))) for _ in range(4)]
    g1 = b"".join(o.g1_to_affine_bytes(p) for p, _ in pairs)
    g2 = b"".join(o.g2_to_affine_bytes(q) for _, q in pairs)
    want_gt = b"".join(o.gt_to_bytes(o.pairing_reference(p, q)) for p, q in pairs)
    a1, a2, b1, b2, want = _check_inputs(rng, 40)
    a1[7] = None; b1[7] = None; want[7] = 1                  # both pairs dead: 1 == 1
    a1[8] = None; want[8] = 0                                # e(inf, Q) = 1 against e(b1, G2) != 1: the live pair alone (1-pair program)
    b1[9] = None; want[9] = 0                                # the other way round
    A1, A2 = b"".join(map(o.g1_to_affine_bytes, a1)), b"".join(map(o.g2_to_affine_bytes, a2))
    B1, B2 = b"".join(map(o.g1_to_affine_bytes, b1)), b"".join(map(o.g2_to_affine_bytes, b2))
    try:
        for coop in (1 << 20, 0):                            # every batch on the cooperative kernel / never
            engine._check(engine.lib.b2k_set_pairing_coop(engine.h, coop))
            assert engine.bls12381_pair(g1, g2) == want_gt, coop
            assert list(engine.bls12381_pairing_check(A1, A2, B1, B2)) == want, coop
        engine._check(engine.lib.b2k_set_pairing_coop(engine.h, 1 << 20))
        bad = bytearray(A1); bad[96 * 3 + 95] ^= 1           # off the curve: that check fails, the others are unaffected
        got = list(engine.bls12381_pairing_check(bytes(bad), A2, B1, B2))
        assert got[3] == 0 and got[:3] == want[:3] and got[4:] == want[4:]
    finally:
        engine._check(engine.lib.b2k_set_pairing_coop(engine.h, 10240))


def test_bls_verify_goes_through_the_cooperative_kernel_for_small_batches(engine):
    """bls.Verify (sign/bls/bls.go:82-96) with n = 1 and n = 5: decompression + hash + the check with the shared G2 generator"""
    from oracle import h2c_bls12381 as h
    rng = random.Random(62)
    for n in (1, 5):
        sks = [rng.randrange(1, o.R) for _ in range(n)]
        msgs = [bytes([i, 7, 9]) * (i + 1) for i in range(n)]
        pks = b"".join(o.g2_compress(o.g2_mul(sk)) for sk in sks)
        sigs = [o.g1_compress(o.g1_mul(sk, h.hash_to_g1(m, h.DST_G1))) for sk, m in zip(sks, msgs)]
        if n > 1:
            sigs[2] = o.g1_compress(o.g1_mul(sks[2] + 1, h.hash_to_g1(msgs[2], h.DST_G1)))      # a wrong signature
        for coop in (10240, 0):
            engine._check(engine.lib.b2k_set_pairing_coop(engine.h, coop))
            ok = engine.bls12381_verify_g1sig(pks, msgs, h.DST_G1, b"".join(sigs))
            assert list(ok) == [0 if (n > 1 and i == 2) else 1 for i in range(n)], (n, coop)
    engine._check(engine.lib.b2k_set_pairing_coop(engine.h, 10240))


@pytest.mark.parametrize("curve", ["bn254", "bn256"])
def test_bn_pairings_on_both_kernels(engine, curve):
    """bn254 / bn256 Pair and ValidatePairing (pairing/bn254/suite.go:133-144, pairing/bn256/suite.go:99-109) through the cooperative
    kernel (small batches, the default) and through the one-per-thread kernel: same bytes as the restatement of the Go source."""
    if curve == "bn254":
        from oracle import bn254 as c, bn254_pairing as b
        pair, check = engine.bn254_pair, engine.bn254_pairing_check
    else:
        from oracle import bn256 as c, bn256_pairing as b
        pair, check = engine.bn256_pair, engine.bn256_pairing_check
    rng = random.Random(63)
    pairs = [(c.G1, b.G2), (None, b.G2), (c.G1, None)] + [(c.g1_mul(rng.randrange(1, c.ORDER)), b.g2_mul(rng.randrange(1, c.ORDER))) for _ in range(3)]
    g1 = b"".join(c.g1_marshal(p) for p, _ in pairs)
    g2 = b"".join(b.g2_marshal(q) for _, q in pairs)
    want_gt = b"".join(b.gt_to_bytes(b.pairing(p, q)) for p, q in pairs)
    x, y = rng.randrange(1, c.ORDER), rng.randrange(1, c.ORDER)
    a1 = [c.g1_mul(x), c.g1_mul(x), None, None, c.g1_mul(x)]
    a2 = [b.g2_mul(y)] * 5
    b1 = [c.g1_mul(x * y % c.ORDER), c.g1_mul((x * y + 1) % c.ORDER), None, c.g1_mul(x), None]
    b2 = [b.G2] * 5
    want = [1, 0, 1, 0, 0]        # right, wrong, both sides 1, 1 against a non-trivial pairing (twice: the live pair alone)
    A1, A2 = b"".join(c.g1_marshal(p) for p in a1), b"".join(b.g2_marshal(q) for q in a2)
    B1, B2 = b"".join(c.g1_marshal(p) for p in b1), b"".join(b.g2_marshal(q) for q in b2)
    try:
        for coop in (1 << 20, 0):
            engine.set_pairing_coop(coop)
            assert pair(g1, g2) == want_gt, (curve, coop)
            assert list(check(A1, A2, B1, B2)) == want, (curve, coop)
    finally:
        engine.set_pairing_coop(10240)


@pytest.mark.parametrize("curve", ["bls12381", "bn254"])
def test_gt_exponentiation_on_both_kernels(engine, curve):
    """GT.Mul (kilic/gt.go:62-71, pairing/bn254/point.go:606-623): a^s for scalars 0, 1, r - 1 and random ones, through the cooperative
    kernel (small batches) and the one-per-thread kernel; a is a pairing value AND an arbitrary Fp12 element (the group law does not
    care, the generic square-and-multiply must not either)."""
    rng = random.Random(64)
    if curve == "bls12381":
        order, one = o.R, o.F12_ONE
        e = o.pairing_reference(o.g1_mul(5), o.g2_mul(7))
        rnd = tuple(tuple((rng.randrange(o.P), rng.randrange(o.P)) for _ in range(3)) for _ in range(2))
        to_bytes, f12_pow = o.gt_to_bytes, o.f12_pow
    else:
        from oracle import bn254 as c4, bn254_pairing as b4
        order = c4.ORDER
        e = b4.pairing(c4.g1_mul(5), b4.g2_mul(7))
        rnd = tuple(tuple((rng.randrange(c4.P), rng.randrange(c4.P)) for _ in range(3)) for _ in range(2))
        to_bytes, f12_pow = b4.gt_to_bytes, b4.f12_pow
    xs = [0, 1, order - 1, rng.randrange(order), rng.randrange(order), 2]
    vals = [e, e, e, e, rnd, rnd]
    sb = b"".join(x.to_bytes(32, "big") for x in xs)
    ab = b"".join(to_bytes(v) for v in vals)
    want = b"".join(to_bytes(f12_pow(v, x)) for v, x in zip(vals, xs))
    try:
        for coop in (1 << 20, 0):
            engine.set_pairing_coop(coop)
            assert engine.gt_exp(curve, sb, ab) == want, (curve, coop)
    finally:
        engine.set_pairing_coop(10240)
