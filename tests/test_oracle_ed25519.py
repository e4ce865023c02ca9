"""CPU tests: edwards25519 Point.Mul oracle, pinned by the reference's seeded KAT (examples/dh_test.go:17-49)
and cross-checked against libsodium; BASELINE configs[0] (batch of 1024 Point.Mul on CPU) as a parity case."""
import hashlib

import pytest

from oracle import ed25519 as ed


def test_reference_seeded_diffie_hellman_kat():
    rng = ed.Blake2Xb(b"")                      # blake2xb.New(nil)
    a = ed.pick_scalar(rng)
    b = ed.pick_scalar(rng)
    A = ed.point_mul(a, None)
    B = ed.point_mul(b, None)
    sa, sb = ed.point_mul(a, B), ed.point_mul(b, A)
    assert sa == sb
    assert sa.hex() == "80ea238cacfdab279626970bba18c69083c7751865dec4c6434bff4351282847"


def test_rfc8032_public_keys():
    # RFC 8032 7.1 test 1 and 2 (sign/eddsa/eddsa_test.go:24-51 uses the same vectors)
    for seed, pub in (("9d61b19deffd5a60ba844af492ec2cc44449c5697b326919703bac031cae7f60",
                       "d75a980182b10ab7d54bfed3c964073a0ee172f3daa62325af021a68f707511a"),
                      ("4ccd089b28ff96da9db6c346ec114e0f5b8a319f35aba624da8cf6ed4fb8a6fb",
                       "3d4017c3e843895a92b70aa74d1b7ebc9c982ccf2ec4968cc0cd55f12af4660c")):
        h = bytearray(hashlib.sha512(bytes.fromhex(seed)).digest()[:32])
        h[0] &= 248; h[31] &= 127; h[31] |= 64
        assert ed.point_mul(bytes(h), None).hex() == pub


def test_batch_1024_point_mul_matches_libsodium():
    nacl = pytest.importorskip("nacl.bindings")
    from kyber_b200 import workload as wl
    n = 1024                                   # BASELINE configs[0]; seed b2k/c1
    s = wl.prng_scalars("b2k/c1", n, ed.L)
    a = wl.prng_scalars("b2k/c1-a", n, ed.L)
    for i in range(0, n, 16):                  # every 16th pair keeps the CPU suite fast; all 1024 under -k full
        k = s[i].to_bytes(32, "little")
        if s[i] == 0 or a[i] == 0:
            continue
        pt = nacl.crypto_scalarmult_ed25519_base_noclamp(a[i].to_bytes(32, "little"))
        assert pt == ed.point_mul(a[i].to_bytes(32, "little"), None)
        assert nacl.crypto_scalarmult_ed25519_noclamp(k, pt) == ed.point_mul(k, pt)


def test_unreduced_scalar_and_torsion_semantics():
    # scalars are raw 256-bit integers (scalar.go:226-233): k and k + l differ on points with a torsion part
    k = 0x123456789abcdef
    t8 = ed.decode(bytes.fromhex("c7176a703d4dd84fba3c0b760d10670f2a2053fa2c39ccc64ec7fd7792ac037a"))   # order-8 point
    assert t8 is not None and ed.scalar_mult(8, t8) == ed.IDENT
    p = ed.add(ed.scalar_mult(77), t8)
    assert ed.scalar_mult(k, p) != ed.scalar_mult(k + ed.L, p)
    assert ed.scalar_mult(k, ed.BASE) == ed.scalar_mult(k + ed.L, ed.BASE)
    assert ed.encode(ed.scalar_mult(0)) == (1).to_bytes(32, "little")
