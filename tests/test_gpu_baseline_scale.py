"""GPU tests at BASELINE.json's full sizes, through size-independent properties (SURVEY 8c tricks):
  configs[1]  BLS12-381 G1 MSM, 2^20 pairs: sum s_i (a_i G) == ((sum s_i a_i) mod r) G
  configs[2]  65 536 independent bls.Verify with 16 corrupted signatures: exactly those fail (mode A);
              BDN-shaped aggregate with 128-bit coefficients over the same keys/signatures verifies (mode B)
  configs[3]  RecoverCommit t = 1024 over bn254 G1 is in tests/test_gpu_bn254.py
Inputs are produced by the engine itself (fixed-base batches) and spot-checked against the oracle."""
import random

import pytest

from kyber_b200 import workload as wl
from oracle import bls12381 as o
from oracle import h2c_bls12381 as h

pytestmark = pytest.mark.gpu


def test_c2_msm_2_pow_20(engine):
    n = 1 << 20
    a = wl.prng_scalars("b2k/c2-a", n, o.R)
    s = wl.prng_scalars("b2k/c2", n, o.R)
    pts = engine.bls12381_g1_mul_batch_affine(wl.scalars_to_bytes(a), wl.G1_BLS12381_AFFINE * n)
    rng = random.Random(1)
    for i in [0, n - 1] + [rng.randrange(n) for _ in range(6)]:
        assert pts[96 * i:96 * i + 96] == o.g1_to_affine_bytes(o.g1_mul(a[i]))
    want = o.g1_compress(o.g1_mul(wl.dot_mod(s, a, o.R)))
    for groups in (1, 4):                      # 4 = the experimental overlapped tail (off by default)
        engine.set_msm_groups(groups)
        try:
            assert engine.bls12381_g1_msm(wl.scalars_to_bytes(s), pts) == want
        finally:
            engine.set_msm_groups(1)
    for rounds in (0, 2):                      # affine pair-tree rounds off / forced (default: automatic, above)
        engine.set_msm_affine(rounds, 0)
        try:
            assert engine.bls12381_g1_msm(wl.scalars_to_bytes(s), pts) == want
        finally:
            engine.set_msm_affine(-1, 0)


def test_c3_65536_signatures_mode_a_and_b(engine):
    n = 1 << 16
    rng = random.Random(3)
    sks = wl.prng_scalars("b2k/c3", n, o.R)
    msg = bytes(range(32))                               # same message for every signer (BDN aggregate needs one message)
    hm = h.hash_to_g1(msg)
    hm_b = o.g1_to_affine_bytes(hm)
    sb = wl.scalars_to_bytes(sks)
    # keys sk_i * G2 and signatures sk_i * H(m), made on the device, compressed on the device
    pk_aff = engine.bls12381_g2_mul_batch_affine(sb, o.g2_to_affine_bytes(o.G2) * n)
    pks = engine.bls12381_g2_mul_batch(b"".join((1).to_bytes(32, "big") for _ in range(n)), pk_aff)
    sig_aff = engine.bls12381_g1_mul_batch_affine(sb, hm_b * n)
    sigs = bytearray(engine.bls12381_g1_mul_batch(b"".join((1).to_bytes(32, "big") for _ in range(n)), sig_aff))
    for i in (0, 12345, n - 1):
        assert pks[96 * i:96 * i + 96] == o.g2_compress(o.g2_mul(sks[i]))
        assert bytes(sigs[48 * i:48 * i + 48]) == o.g1_compress(o.g1_mul(sks[i], hm))
    # ---- mode A: corrupt 16 signatures (replace by another signer's), exactly those must fail
    bad = sorted(rng.sample(range(n), 16))
    good_sigs = bytes(sigs)
    for i in bad:
        j = (i + 1) % n
        sigs[48 * i:48 * i + 48] = good_sigs[48 * j:48 * j + 48]
    ok = engine.bls12381_verify_g1sig(pks, [msg] * n, h.DST_G1, bytes(sigs))
    assert [i for i in range(n) if ok[i] == 0] == bad
    # ---- mode B: BDN-shaped aggregate (sign/bdn/bdn.go:126-181) with 128-bit coefficients c_i + 1
    coefs = [rng.randrange(1 << 128) + 1 for _ in range(n)]
    cb = wl.scalars_to_bytes(coefs)
    agg_sig = engine.bls12381_g1_msm(cb, sig_aff)                        # sum (c_i+1) S_i        (48 B)
    agg_key = engine.bls12381_g2_msm(cb, pk_aff)                         # sum (c_i+1) PK_i       (96 B)
    assert agg_sig == o.g1_compress(o.g1_mul(wl.dot_mod(coefs, sks, o.R), hm))
    assert agg_key == o.g2_compress(o.g2_mul(wl.dot_mod(coefs, sks, o.R)))
    assert engine.bls12381_verify_g1sig(agg_key, [msg], h.DST_G1, agg_sig) == b"\x01"      # one bls.Verify of the aggregate
