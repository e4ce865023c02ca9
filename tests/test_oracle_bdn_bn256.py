"""CPU tests: pin oracle/bn256.py and oracle/bdn.py on the reference's byte-exact BDN fixtures
(sign/bdn/bdn_vartime_test.go:24-48, :90-135; data in tests/golden/bdn_bn256_fixtures.json)."""
import json
import os

from oracle import bdn, bn256 as o

FX = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bdn_bn256_fixtures.json")))


def test_hash_point_to_r_and_aggregate_key():
    f = FX["hash_point_to_r"]
    pubs = [o.g2_mul(k) for k in (1, 2, 3)]
    coefs = bdn.hash_point_to_r([o.g2_marshal(p) for p in pubs], o.ORDER)
    assert ["%x" % c for c in coefs] == f["coefs"]
    agg = None
    for c, p in zip(coefs, pubs):
        agg = o.g2_add(agg, o.g2_add(o.g2_mul(c, p), p))
    assert o.g2_marshal(agg).hex() == f["agg_key"]


def test_bdn_fixtures_signatures_and_aggregates():
    f = FX["fixtures"]
    msg = f["msg"].encode()
    privs = [int(x, 16) for x in f["private"]]
    pubs = [o.g2_unmarshal(bytes.fromhex(x)) for x in f["public"]]
    for sk, pk in zip(privs, pubs):
        assert o.g2_mul(sk) == pk
    hm = o.hash_to_g1(msg)
    sigs = [o.g1_mul(sk, hm) for sk in privs]
    assert [o.g1_marshal(s).hex() for s in sigs] == f["sig"]
    coefs = bdn.hash_point_to_r([bytes.fromhex(x) for x in f["public"]], o.ORDER)
    agg_sig, agg_key = None, None
    for i in f["mask_enabled"]:
        agg_sig = o.g1_add(agg_sig, o.g1_add(o.g1_mul(coefs[i], sigs[i]), sigs[i]))
        agg_key = o.g2_add(agg_key, o.g2_add(o.g2_mul(coefs[i], pubs[i]), pubs[i]))
    assert o.g1_marshal(agg_sig).hex() == f["agg_sig"]
    assert o.g2_marshal(agg_key).hex() == f["agg_key"]
