

def test_bn256_hashg1_oracle_reproduces_the_reference_kats():
    """pairing/bn256/hash_test.go:11-57: HashG1([]byte{byte(i)}, nil) for i = 0..10 -- the oracle's HKDF + SvdW restatement is
    PINNED by these 11 marshalled points (tests/golden/bn256_hashg1_vectors.json, extracted by make_golden.py)"""
    import json, os
    from oracle import bn256 as c6
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bn256_hashg1_vectors.json")))["cases"]
    assert len(fx) == 11
    for c in fx:
        pt = c6.hash_g1(bytes.fromhex(c["msg_hex"]), None)
        assert c6.g1_marshal(pt).hex() == c["point"]
        assert (pt[1] ** 2 - pt[0] ** 3 - c6.B) % c6.P == 0


def test_recover_pubpoly_oracle_returns_the_dealers_commitments():
    """share.RecoverPubPoly restated (share/poly.go:480-545): from t shares f(i+1) G of a random polynomial the commitments f_k G
    come back; commits[0] is RecoverCommit's point"""
    import random
    from oracle import bn254 as c4, share_poly
    rng = random.Random(9)
    t = 5
    coeffs = [rng.randrange(c4.ORDER) for _ in range(t)]
    idx = [0, 2, 3, 7, 11, 12]
    shares = [(i, c4.g1_mul(sum(c * pow(i + 1, k, c4.ORDER) for k, c in enumerate(coeffs)) % c4.ORDER)) for i in idx]
    commits = share_poly.recover_pubpoly(c4, shares, t)
    assert commits == [c4.g1_mul(c) for c in coeffs]
    assert commits[0] == share_poly.recover_commit(c4, shares, t)
