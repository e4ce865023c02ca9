"""GPU parity: BLS12-381 pairings through the C ABI vs the Python oracle (GT bytes with the exact
exponent, pairing-equation booleans).  Mirrors the reference's pairing property tests
(pairing/bls12381/bls12381_test.go:448-474 bilinearity, :580-631 product identity) and
Suite.ValidatePairing semantics (kilic/suite.go:57-68)."""
import random

import pytest

from oracle import bls12381 as o

pytestmark = pytest.mark.gpu


def test_pair_gt_bytes_match_oracle(engine):
    rng = random.Random(21)
    pairs = [(o.G1, o.G2)]
    for _ in range(5):
        pairs.append((o.g1_mul(rng.randrange(1, o.R)), o.g2_mul(rng.randrange(1, o.R))))
    pairs.append((None, o.G2))                  # infinity operands -> GT identity
    pairs.append((o.G1, None))
    g1 = b"".join(o.g1_to_affine_bytes(p) for p, _ in pairs)
    g2 = b"".join(o.g2_to_affine_bytes(q) for _, q in pairs)
    gt = engine.bls12381_pair(g1, g2)
    for i, (p, q) in enumerate(pairs):
        assert gt[576 * i:576 * (i + 1)] == o.gt_to_bytes(o.pairing(p, q)), i


def test_pairing_bilinearity_on_device(engine):
    rng = random.Random(22)
    a, b = rng.randrange(1, o.R), rng.randrange(1, o.R)
    # e(aG1, bG2) == e(abG1, G2) == e(G1, abG2)
    g1 = b"".join(o.g1_to_affine_bytes(p) for p in (o.g1_mul(a), o.g1_mul(a * b % o.R), o.G1))
    g2 = b"".join(o.g2_to_affine_bytes(q) for q in (o.g2_mul(b), o.G2, o.g2_mul(a * b % o.R)))
    gt = engine.bls12381_pair(g1, g2)
    assert gt[:576] == gt[576:1152] == gt[1152:]


def test_pairing_check_batch(engine):
    rng = random.Random(23)
    n = 70
    a1, a2, b1, b2, want = [], [], [], [], []
    for i in range(n):
        x, y = rng.randrange(1, o.R), rng.randrange(1, o.R)
        good = (i % 5 != 3)
        a1.append(o.g1_mul(x)); a2.append(o.g2_mul(y))
        b1.append(o.g1_mul(x * y % o.R if good else (x * y + 1) % o.R)); b2.append(o.G2)
        want.append(1 if good else 0)
    ok = engine.bls12381_pairing_check(b"".join(map(o.g1_to_affine_bytes, a1)), b"".join(map(o.g2_to_affine_bytes, a2)),
                                       b"".join(map(o.g1_to_affine_bytes, b1)), b"".join(map(o.g2_to_affine_bytes, b2)))
    assert list(ok) == want


def test_bls_verify_equation_sigs_on_g1(engine):
    """bls.Verify (sign/bls/bls.go:82-96) with signatures on G1: e(H(m), X) == e(sig, G2base);
    H(m) is stood in by a known multiple of the generator here (hash-to-curve is tested separately)."""
    rng = random.Random(24)
    sk, h = rng.randrange(1, o.R), rng.randrange(1, o.R)
    hm, pk = o.g1_mul(h), o.g2_mul(sk)
    sig = o.g1_mul(sk, hm)
    args = [o.g1_to_affine_bytes(hm), o.g2_to_affine_bytes(pk), o.g1_to_affine_bytes(sig), o.g2_to_affine_bytes(o.G2)]
    assert engine.bls12381_pairing_check(*args) == b"\x01"
    args[2] = o.g1_to_affine_bytes(o.g1_mul(sk + 1, hm))
    assert engine.bls12381_pairing_check(*args) == b"\x00"
