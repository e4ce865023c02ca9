"""GPU parity: batched hash-to-G1 and batched bls.Verify (signatures on G1) through the C ABI, pinned by
the reference's own KATs: TestSignatureEdgeCase (pairing/bls12381/bls12381_test.go:877-904) and the drand
vectors of pairing/bls12381/kilic/suite_test.go:17-46,84-106 (tests/golden/bls12381_signature_kats.json)."""
import hashlib
import json
import os
import random

import pytest

from oracle import bls12381 as o
from oracle import h2c_bls12381 as h

pytestmark = pytest.mark.gpu
KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bls12381_signature_kats.json")))


def test_hash_to_g1_matches_oracle(engine):
    rng = random.Random(51)
    msgs = [b"", b"abc", bytes(range(32)), rng.randbytes(55), rng.randbytes(56), rng.randbytes(64), rng.randbytes(119),
            rng.randbytes(200), rng.randbytes(1)]
    for dst in (h.DST_G1, h.DST_G2, b"b2k-test"):
        out = engine.bls12381_hash_to_g1(msgs, dst)
        for i, m in enumerate(msgs):
            assert out[96 * i:96 * i + 96] == o.g1_to_affine_bytes(h.hash_to_g1(m, dst)), (i, dst)


def test_reference_kats_verify_on_device(engine):
    e = KAT["edge_case_g1"]
    d = KAT["sig_on_g1_g2domain"]
    drand_msg = hashlib.sha256((d["round"]).to_bytes(8, "big")).digest()
    # TestSignatureEdgeCase: default G1 DST
    ok = engine.bls12381_verify_g1sig(bytes.fromhex(e["pk_g2"]), [bytes.fromhex(e["msg"])], h.DST_G1, bytes.fromhex(e["sig_g1"]))
    assert ok == b"\x01"
    # drand vector: must FAIL with the G1 DST and PASS with the G2 DST (kilic/suite_test.go:17-46)
    pk, sig = bytes.fromhex(d["pk_g2"]), bytes.fromhex(d["sig_g1"])
    assert engine.bls12381_verify_g1sig(pk, [drand_msg], h.DST_G1, sig) == b"\x00"
    assert engine.bls12381_verify_g1sig(pk, [drand_msg], h.DST_G2, sig) == b"\x01"


def test_verify_batch_with_corrupted_entries(engine):
    """BASELINE configs[2] mode A in miniature: independent verifications, the corrupted ones must be exactly
    the ones that fail (wrong signature, wrong message, signature outside the subgroup, malformed key)."""
    rng = random.Random(52)
    n = 24
    sks = [rng.randrange(1, o.R) for _ in range(n)]
    msgs = [rng.randbytes(32) for _ in range(n)]
    pks = [o.g2_compress(o.g2_mul(sk)) for sk in sks]
    sigs = [o.g1_compress(o.g1_mul(sk, h.hash_to_g1(m))) for sk, m in zip(sks, msgs)]
    want = [1] * n
    sigs[3] = o.g1_compress(o.g1_mul(sks[3] + 1, h.hash_to_g1(msgs[3]))); want[3] = 0       # wrong signature
    msgs[7] = msgs[7][:-1] + bytes([msgs[7][-1] ^ 1]); want[7] = 0                            # wrong message
    while True:                                                                                # sig not in G1
        x = rng.randrange(o.P)
        y = o.fp_sqrt((x ** 3 + 4) % o.P)
        if y is not None:
            break
    sigs[11] = o.g1_compress((x, y)); want[11] = 0
    pks[13] = bytes([pks[13][0] & 0x7F]) + pks[13][1:]; want[13] = 0                         # compression flag cleared
    pks[17], want[17] = pks[18], 0                                                            # someone else's key
    ok = engine.bls12381_verify_g1sig(b"".join(pks), msgs, h.DST_G1, b"".join(sigs))
    assert list(ok) == want


def test_hash_to_g2_and_drand_kat_on_device(engine):
    """signatures on G2 (the drand default): hash-to-G2 parity and the reference's KAT kilic/suite_test.go:48-72."""
    from oracle import h2c_bls12381_g2 as h2
    rng = random.Random(53)
    msgs = [b"", b"abc", rng.randbytes(32), rng.randbytes(100)]
    out = engine.bls12381_hash_to_g2(msgs, h.DST_G2)
    for i, m in enumerate(msgs):
        assert out[192 * i:192 * i + 192] == o.g2_to_affine_bytes(h2.hash_to_g2(m, h.DST_G2)), i
    k = KAT["sig_on_g2"]
    msg = hashlib.sha256(bytes.fromhex(k["prev_sig"]) + (k["round"]).to_bytes(8, "big")).digest()
    pk, sig = bytes.fromhex(k["pk_g1"]), bytes.fromhex(k["sig_g2"])
    assert engine.bls12381_verify_g2sig(pk, [msg], h.DST_G2, sig) == b"\x01"
    assert engine.bls12381_verify_g2sig(pk, [msg], h.DST_G1, sig) == b"\x00"          # wrong domain
    bad = msg[:-1] + bytes([msg[-1] ^ 1])
    assert engine.bls12381_verify_g2sig(pk * 2, [msg, bad], h.DST_G2, sig * 2) == b"\x01\x00"
