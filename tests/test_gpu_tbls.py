"""GPU parity for the threshold-BLS path: share.PubPoly.Eval for many indices
(share/poly.go:340-357), share.RecoverCommit on BLS12-381 G1/G2 (share/poly.go:449-476) and the whole of
tbls.Recover (sign/tbls/tbls.go:118-151): evaluate the public polynomial at every signer index, verify every partial
signature, recover the group signature by Lagrange interpolation -- and the result must be the ordinary BLS signature
of the group secret."""
import random

import pytest

from kyber_b200 import workload as wl
from oracle import bls12381 as o
from oracle import h2c_bls12381 as h
from oracle import share_poly

pytestmark = pytest.mark.gpu


class _G1:
    ORDER = o.R
    g1_add = staticmethod(o.g1_add)
    g1_mul = staticmethod(o.g1_mul)


def _poly_eval(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % o.R
    return acc


def test_pubpoly_eval_g1_and_g2(engine):
    rng = random.Random(81)
    t = 9
    coeffs = [rng.randrange(o.R) for _ in range(t)]
    idx = [0, 1, 5, 77, 4095, 2 ** 31]
    c1 = b"".join(o.g1_to_affine_bytes(o.g1_mul(c)) for c in coeffs)           # PriPoly.Commit: C_j = a_j * G
    out = engine.bls12381_pubpoly_eval(1, c1, idx)
    for k, i in enumerate(idx):
        assert out[96 * k:96 * k + 96] == o.g1_to_affine_bytes(o.g1_mul(_poly_eval(coeffs, i + 1))), i
    c2 = b"".join(o.g2_to_affine_bytes(o.g2_mul(c)) for c in coeffs)
    out = engine.bls12381_pubpoly_eval(2, c2, idx[:4])
    for k, i in enumerate(idx[:4]):
        assert out[192 * k:192 * k + 192] == o.g2_to_affine_bytes(o.g2_mul(_poly_eval(coeffs, i + 1))), i


def test_recover_commit_bls12381(engine):
    rng = random.Random(82)
    t = 40
    coeffs = [rng.randrange(o.R) for _ in range(t)]
    idx = sorted(rng.sample(range(200), t))
    ys = [_poly_eval(coeffs, i + 1) for i in idx]
    p1 = engine.bls12381_g1_mul_batch_affine(wl.scalars_to_bytes(ys), wl.G1_BLS12381_AFFINE * t)
    assert engine.bls12381_recover_commit(1, idx, p1) == o.g1_compress(o.g1_mul(coeffs[0]))
    shares = [(i, o.g1_from_affine_bytes(p1[96 * k:96 * k + 96])) for k, i in enumerate(idx)]
    assert engine.bls12381_recover_commit(1, idx, p1) == o.g1_compress(share_poly.recover_commit(_G1, shares, t))
    p2 = engine.bls12381_g2_mul_batch_affine(wl.scalars_to_bytes(ys), o.g2_to_affine_bytes(o.G2) * t)
    assert engine.bls12381_recover_commit(2, idx, p2) == o.g2_compress(o.g2_mul(coeffs[0]))


def test_pubpoly_check_batch_of_deals(engine):
    """the share-verification loops of share/vss and share/dkg (PubPoly.Check, share/poly.go:405-409,
    once per deal in vss.go:636-645 / dkg.go:489-494) as one batch: m dealers, each with its own commitment polynomial,
    n shares each; corrupted shares, a share for the wrong index and an out-of-range share must be the only failures."""
    rng = random.Random(84)
    m, t, n = 5, 7, 11
    from oracle import bn254 as o4
    for curve, order, commit in (("bls12381_g1", o.R, lambda c: o.g1_to_affine_bytes(o.g1_mul(c))),
                                 ("bls12381_g2", o.R, lambda c: o.g2_to_affine_bytes(o.g2_mul(c))),
                                 ("bn254", o4.ORDER, lambda c: o4.g1_marshal(o4.g1_mul(c)))):
        polys = [[rng.randrange(order) for _ in range(t)] for _ in range(m)]
        polys[3][0] = 0                                            # a zero coefficient: commitment at infinity
        idx = [[rng.randrange(0, 1000) for _ in range(n)] for _ in range(m)]
        idx[0][0] = 0
        idx[1][1] = 2 ** 32 - 1                                     # x = 2^32
        def ev(cs, x):
            acc = 0
            for c in reversed(cs):
                acc = (acc * x + c) % order
            return acc
        sh = [[ev(polys[d], idx[d][k] + 1) for k in range(n)] for d in range(m)]
        bad = {(0, 3), (2, 0), (4, 10), (1, 5), (3, 7)}
        sh[0][3] = (sh[0][3] + 1) % order                           # corrupted share
        sh[2][0] = ev(polys[1], idx[2][0] + 1)                      # share of another dealer's polynomial
        sh[4][10] = ev(polys[4], idx[4][10] + 2)                    # share for the wrong index
        raw = [[v.to_bytes(32, "big") for v in row] for row in sh]
        raw[1][5] = order.to_bytes(32, "big")                       # not below the group order
        raw[3][7] = bytes(32) if sh[3][7] else (1).to_bytes(32, "big")
        commits = b"".join(commit(c) for cs in polys for c in cs)
        ok = engine.pubpoly_check(curve, commits, t, idx, b"".join(b for row in raw for b in row))
        for d in range(m):
            for k in range(n):
                assert ok[d * n + k] == (0 if (d, k) in bad else 1), (curve, d, k)


def test_tbls_recover_flow(engine):
    """t-of-n threshold BLS, signatures on G1 / keys on G2: all heavy steps on the engine."""
    rng = random.Random(83)
    t, n = 16, 24
    coeffs = [rng.randrange(o.R) for _ in range(t)]                 # secret polynomial; group secret = coeffs[0]
    msg = b"threshold message"
    hm = h.hash_to_g1(msg)
    pub_commits = b"".join(o.g2_to_affine_bytes(o.g2_mul(c)) for c in coeffs)     # PubPoly on G2
    signers = list(range(n))
    shares = [_poly_eval(coeffs, i + 1) for i in signers]
    part = engine.bls12381_g1_mul_batch(wl.scalars_to_bytes(shares), o.g1_to_affine_bytes(hm) * n)   # partial sigs (48 B)
    part = bytearray(part)
    part[48 * 3:48 * 4] = part[48 * 4:48 * 5]                        # signer 3 sends a wrong partial signature
    # tbls.Recover: public.Eval(idx) for every share (n evaluations of a degree t-1 commitment polynomial)
    pk_aff = engine.bls12381_pubpoly_eval(2, pub_commits, signers)
    pk_c = engine.bls12381_g2_mul_batch(b"".join((1).to_bytes(32, "big") for _ in range(n)), pk_aff)   # compress
    ok = engine.bls12381_verify_g1sig(pk_c, [msg] * n, h.DST_G1, bytes(part))
    assert list(ok) == [0 if i == 3 else 1 for i in range(n)]
    good = [i for i in signers if ok[i]][:t]
    pts, okd = engine.bls12381_g1_decompress(b"".join(bytes(part[48 * i:48 * i + 48]) for i in good))
    assert set(okd) == {1}
    sig = engine.bls12381_recover_commit(1, good, pts)
    assert sig == o.g1_compress(o.g1_mul(coeffs[0], hm))             # = bls.Sign(group secret, msg)
    group_pk = o.g2_compress(o.g2_mul(coeffs[0]))
    assert engine.bls12381_verify_g1sig(group_pk, [msg], h.DST_G1, sig) == b"\x01"


def test_commit_batch_and_recover_pubpoly(engine):
    """PriPoly.Commit as a fixed-base batch (share/poly.go:143-149) and share.RecoverPubPoly (poly.go:480-508): from t public
    shares of a random polynomial the engine must return exactly the dealer's commitments a_k * G (and the oracle's
    restatement of lagrangeBasis agrees); G1, G2 and bn254; arbitrary base point; duplicate index refused."""
    from kyber_b200 import B2KError
    from oracle import bn254 as c4
    rng = random.Random(85)
    t = 17
    coeffs = [rng.randrange(o.R) for _ in range(t)]
    sc = b"".join(o.scalar_to_bytes(c) for c in coeffs)
    commits = engine.commit_batch("bls12381_g1", sc)                             # a_k * G1
    assert commits == b"".join(o.g1_to_affine_bytes(o.g1_mul(c)) for c in coeffs)
    H = o.g1_mul(rng.randrange(1, o.R))
    assert engine.commit_batch("bls12381_g1", sc, o.g1_to_affine_bytes(H)) == b"".join(o.g1_to_affine_bytes(o.g1_mul(c, H)) for c in coeffs)
    idx = sorted(rng.sample(range(100), t))
    shares = engine.bls12381_pubpoly_eval(1, commits, idx)                       # PubPoly.Eval at the share indices
    got = engine.recover_pubpoly("bls12381_g1", idx, shares)
    assert got == commits
    # the oracle's restatement of RecoverPubPoly on a small case
    small = [(i, o.g1_from_affine_bytes(shares[96 * k:96 * k + 96])) for k, i in enumerate(idx[:5])]
    want = share_poly.recover_pubpoly(_G1, small, 5)
    assert engine.recover_pubpoly("bls12381_g1", idx[:5], shares[:5 * 96]) == b"".join(o.g1_to_affine_bytes(p) for p in want)
    # commits[0] is RecoverCommit's point
    assert o.g1_compress(o.g1_from_affine_bytes(got[:96])) == engine.bls12381_recover_commit(1, idx, shares)
    # G2
    c2 = engine.commit_batch("bls12381_g2", sc[:32 * 6])
    assert c2 == b"".join(o.g2_to_affine_bytes(o.g2_mul(c)) for c in coeffs[:6])
    sh2 = engine.bls12381_pubpoly_eval(2, c2, idx[:6])
    assert engine.recover_pubpoly("bls12381_g2", idx[:6], sh2) == c2
    # bn254
    co4 = [rng.randrange(c4.ORDER) for _ in range(9)]
    cm4 = engine.commit_batch("bn254", b"".join(c.to_bytes(32, "big") for c in co4))
    assert cm4 == b"".join(c4.g1_marshal(c4.g1_mul(c)) for c in co4)
    i4 = [3, 4, 9, 10, 11, 40, 41, 77, 500]
    sh4 = b"".join(c4.g1_marshal(c4.g1_mul(sum(c * pow(i + 1, k, c4.ORDER) for k, c in enumerate(co4)) % c4.ORDER)) for i in i4)
    assert engine.recover_pubpoly("bn254", i4, sh4) == cm4
    with pytest.raises(B2KError):
        engine.recover_pubpoly("bls12381_g1", [1, 2, 2], shares[:3 * 96])


def test_recover_pubpoly_large_threshold(engine):
    """t = 300: the master-polynomial kernel runs 300 block-synchronised steps; result = the dealer's commitments"""
    rng = random.Random(86)
    t = 300
    coeffs = [rng.randrange(o.R) for _ in range(t)]
    commits = engine.commit_batch("bls12381_g1", b"".join(o.scalar_to_bytes(c) for c in coeffs))
    idx = sorted(rng.sample(range(2000), t))
    shares = engine.bls12381_pubpoly_eval(1, commits, idx)
    assert engine.recover_pubpoly("bls12381_g1", idx, shares) == commits
    for k in (0, 1, t - 1):
        assert commits[96 * k:96 * k + 96] == o.g1_to_affine_bytes(o.g1_mul(coeffs[k]))
