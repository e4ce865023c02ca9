"""GPU parity on the reference's ONLY byte-exact scalar-mul / MSM vectors on a pairing curve: the BDN fixtures on
bn256 (sign/bdn/bdn_vartime_test.go:24-48, :90-135).  Signatures = x*H(m) through mul_batch, aggregate signature
and aggregate key through the MSM with BDN coefficients (c_i + 1); every byte must equal the Go fixture."""
import json
import os

import pytest

from oracle import bdn, bn256 as o

pytestmark = pytest.mark.gpu
FX = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bdn_bn256_fixtures.json")))


def _sc(vals):
    return b"".join((v % o.ORDER).to_bytes(32, "big") for v in vals)


def test_signatures_are_the_fixture_bytes(engine):
    f = FX["fixtures"]
    hm = engine.bn256_hash_to_g1([f["msg"].encode()])         # device: SHA-256 try-and-increment (bn256/point.go:261-312)
    assert hm == o.g1_marshal(o.hash_to_g1(f["msg"].encode()))
    privs = [int(x, 16) for x in f["private"]]
    out = engine.bn256_g1_mul_batch(_sc(privs), hm * 3)       # sig_i = x_i * H(m)   (bls.Sign, sign/bls/bls.go:67-80)
    assert [out[64 * i:64 * i + 64].hex() for i in range(3)] == f["sig"]
    base2 = o.g2_marshal(o.G2)
    pk = engine.bn256_g2_mul_batch(_sc(privs), base2 * 3)     # public keys x_i * G2
    assert [pk[128 * i:128 * i + 128].hex() for i in range(3)] == f["public"]


def test_aggregate_signature_and_key_are_the_fixture_bytes(engine):
    f = FX["fixtures"]
    pub_bytes = [bytes.fromhex(x) for x in f["public"]]
    coefs = bdn.hash_point_to_r(pub_bytes, o.ORDER)           # oracle (checker)
    en = f["mask_enabled"]
    fac = engine.bdn_coefficients(b"".join(pub_bytes), 128, add_one=True)   # product: host BLAKE2Xs inside the library
    scal = b"".join(fac[32 * i:32 * i + 32] for i in en)      # c_i * S_i + S_i = (c_i + 1) * S_i
    assert scal == _sc([coefs[i] + 1 for i in en])
    sigs = b"".join(bytes.fromhex(f["sig"][i]) for i in en)
    assert engine.bn256_g1_msm(scal, sigs).hex() == f["agg_sig"]
    keys = b"".join(pub_bytes[i] for i in en)
    assert engine.bn256_g2_msm(scal, keys).hex() == f["agg_key"]
    for c in (4, 9, 16):
        engine.set_msm_window(c)
        try:
            assert engine.bn256_g1_msm(scal, sigs).hex() == f["agg_sig"]
        finally:
            engine.set_msm_window(0)


def test_hash_point_to_r_reference_vector(engine):
    f = FX["hash_point_to_r"]
    base2 = o.g2_marshal(o.G2)
    pubs = engine.bn256_g2_mul_batch(_sc([1, 2, 3]), base2 * 3)
    pub_bytes = [pubs[128 * i:128 * i + 128] for i in range(3)]
    coefs = bdn.hash_point_to_r(pub_bytes, o.ORDER)
    assert ["%x" % c for c in coefs] == f["coefs"]
    assert engine.bdn_coefficients(pubs, 128) == _sc(coefs)
    assert engine.bn256_g2_msm(engine.bdn_coefficients(pubs, 128, add_one=True), pubs).hex() == f["agg_key"]


def test_bn256_pairing_and_full_bdn_verification(engine):
    """bn256 pairing on the device: GT bytes vs the restatement of pairing/bn256/optate.go, then the whole BDN flow of
    the fixtures: every signature verifies (bls.Verify: e(H(m), X) == e(sig, G2 base), sign/bls/bls.go:36-38,82-96) and
    so does the aggregate signature under the aggregate key."""
    from oracle import bn256_pairing as bp
    f = FX["fixtures"]
    gt = engine.bn256_pair(o.g1_marshal(o.G1) + o.g1_marshal(o.g1_mul(7)), o.g2_marshal(o.G2) + o.g2_marshal(o.g2_mul(11)))
    assert gt[:384] == bp.gt_to_bytes(bp.pairing(o.G1, o.G2))
    assert gt[384:] == bp.gt_to_bytes(bp.pairing(o.g1_mul(7), o.g2_mul(11)))
    hm = engine.bn256_hash_to_g1([f["msg"].encode()])
    base2 = o.g2_marshal(o.G2)
    sigs = [bytes.fromhex(x) for x in f["sig"]] + [bytes.fromhex(f["agg_sig"]), bytes.fromhex(f["sig"][0])]
    keys = [bytes.fromhex(x) for x in f["public"]] + [bytes.fromhex(f["agg_key"]), bytes.fromhex(f["public"][1])]
    n = len(sigs)
    ok = engine.bn256_pairing_check(hm * n, b"".join(keys), b"".join(sigs), base2 * n)
    assert list(ok) == [1, 1, 1, 1, 0]          # three signatures, the aggregate, and a signature under the wrong key
