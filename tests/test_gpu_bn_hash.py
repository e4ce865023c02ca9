"""GPU parity: hash-to-G1 on the BN curves through the C ABI.  bn254 against the reference's own known answers
(pairing/bn254/point_test.go:14-47) and the oracle on ragged message lengths; bn256 against the oracle (pinned by the
byte-exact BDN signature fixtures); then bls.Sign / bls.Verify on bn254 (sign/bls/bls.go:67-96) with every step on the
device: hash, x * H(m), pairing check."""
import json
import os
import random

import pytest

from oracle import bn254 as o4
from oracle import bn254_pairing as bp4
from oracle import bn254_hash as bh
from oracle import bn256 as o6

pytestmark = pytest.mark.gpu
FX = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bn254_hash_vectors.json")))


def test_bn254_hash_to_point_reference_vectors(engine):
    dst = FX["hash_to_point"]["dst"].encode()
    msgs = [bytes.fromhex(c["msg_hex"]) for c in FX["hash_to_point"]["cases"]]
    out = engine.bn254_hash_to_g1(msgs, dst)
    assert [out[64 * i:64 * i + 64].hex() for i in range(len(msgs))] == [c["point"] for c in FX["hash_to_point"]["cases"]]


def test_bn254_hash_ragged_batch_matches_oracle(engine):
    rng = random.Random(31)
    msgs = [b"", b"a", bytes(135), bytes(136), bytes(137), bytes(rng.getrandbits(8) for _ in range(300))]
    msgs += [bytes(rng.getrandbits(8) for _ in range(rng.randrange(1, 80))) for _ in range(150)]
    dst = b"BN254G1_XMD:KECCAK-256_SVDW_RO_"
    out = engine.bn254_hash_to_g1(msgs)
    for i, m in enumerate(msgs):
        assert out[64 * i:64 * i + 64] == o4.g1_marshal(bh.hash_to_g1(dst, m)), i


def test_bn256_hash_matches_oracle(engine):
    rng = random.Random(32)
    msgs = [b"", b"abc", bytes(55), bytes(56), bytes(64)] + [bytes(rng.getrandbits(8) for _ in range(rng.randrange(1, 120))) for _ in range(200)]
    out = engine.bn256_hash_to_g1(msgs)
    for i, m in enumerate(msgs):
        assert out[64 * i:64 * i + 64] == o6.g1_marshal(o6.hash_to_g1(m)), i


def test_bn256_hashg1_reference_kats_and_oracle(engine):
    """bn256 HashG1 (pairing/bn256/hash.go:10-110) on the device: the reference's 11 marshalled points (hash_test.go:11-57,
    msg = one byte i, nil dst), then ragged messages with and without a dst against the oracle"""
    import json, os
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bn256_hashg1_vectors.json")))["cases"]
    out = engine.bn256_hash_g1([bytes.fromhex(c["msg_hex"]) for c in fx])
    assert out.hex() == "".join(c["point"] for c in fx)
    rng = random.Random(34)
    msgs = [b"", bytes(55), bytes(64), bytes(119), bytes(120)] + [bytes(rng.getrandbits(8) for _ in range(rng.randrange(1, 200))) for _ in range(120)]
    for dst in (b"", b"B2K-DST", bytes(70)):
        out = engine.bn256_hash_g1(msgs, dst)
        for i, m in enumerate(msgs):
            assert out[64 * i:64 * i + 64] == o6.g1_marshal(o6.hash_g1(m, dst or None)), (i, dst)


def test_bls_sign_verify_on_bn254_all_on_device(engine):
    rng = random.Random(33)
    n = 12
    msgs = [b"round %d" % i for i in range(n)]
    xs = [rng.randrange(1, o4.ORDER) for _ in range(n)]
    sc = b"".join(x.to_bytes(32, "big") for x in xs)
    hm = engine.bn254_hash_to_g1(msgs)
    sigs = engine.bn254_g1_mul_batch(sc, hm)                                   # sig = x * H(m)
    base2 = bp4.g2_marshal(bp4.G2)
    pks = engine.bn254_g2_mul_batch(sc, base2 * n)                             # X = x * G2
    for i in (0, 5):
        assert sigs[64 * i:64 * i + 64] == o4.g1_marshal(o4.g1_mul(xs[i], bh.hash_to_g1(b"BN254G1_XMD:KECCAK-256_SVDW_RO_", msgs[i])))
    ok = engine.bn254_pairing_check(hm, pks, sigs, base2 * n)                  # e(H(m), X) == e(sig, G2)
    assert list(ok) == [1] * n
    bad = bytearray(pks)
    bad[128:256] = pks[:128]                                                   # key of signer 0 for signature 1
    ok = engine.bn254_pairing_check(hm, bytes(bad), sigs, base2 * n)
    assert list(ok) == [1, 0] + [1] * (n - 2)
