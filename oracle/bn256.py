"""CPU ORACLE (test infrastructure, NOT product code) -- bn256 (pairing/bn256) G1, G2 and Hash in Python.

Restates the in-tree reference:
  constants     pairing/bn256/constants.go:16-22 (u, p, Order), curve.go:16-24 (b = 3, generator (1,-2)),
                twist.go:22-33 (twistGen, Montgomery limbs R = 2^256), constants.go twistB = 3/xi, xi = i+3
  gfP2          pairing/bn256/gfp2.go:13-15: struct {x, y} = x*i + y, i^2 = -1
  Mul           curve.go:189-203, twist.go:162-175 (the affine result is what MarshalBinary shows)
  MarshalBinary point.go:170-192 (G1: x||y), :423-452 (G2: x.imag||x.real||y.imag||y.real), infinity = zeros
  Hash          point.go:261-312: x = SHA-256(m) mod p, increment until x^3+3 is a square, y = (x^3+3)^((p+1)/4)
                (the root big.Int.ModSqrt returns for p = 3 mod 4; no sign normalisation)
Pinned by the reference's BDN fixtures (sign/bdn/bdn_vartime_test.go:24-48, :90-135) in tests/.
"""
from __future__ import annotations
import hashlib

U = 6518589491078791937
P = 36 * U ** 4 + 36 * U ** 3 + 24 * U ** 2 + 6 * U + 1
ORDER = 36 * U ** 4 + 36 * U ** 3 + 18 * U ** 2 + 6 * U + 1
assert P == 65000549695646603732796438742359905742825358107623003571877145026864184071783
B = 3
G1 = (1, P - 2)


def _limbs(*ws):  # little-endian 64-bit limbs, Montgomery form with R = 2^256
    v = sum(w << (64 * i) for i, w in enumerate(ws))
    return v * pow(1 << 256, -1, P) % P


# Fp2 elements are (real, imag)
G2 = ((_limbs(0x88f9f11da7cdc184, 0x18293f95d69509d3, 0xb5ce0c55a735d5a1, 0x15134189bfd45a0),
       _limbs(0x402c4ab7139e1404, 0xce1c368a183d85a4, 0xd67cf9a6cb8d3983, 0x3cf246bbc2a9fbe8)),
      (_limbs(0xc2e07c1463ea9e56, 0xee4442052072ebd2, 0x561a519486036937, 0x5bd9394cc0d2cce),
       _limbs(0xbfac7d731e9e87a2, 0xa50bb8007962e441, 0xafe910a4e8270556, 0x5075c5429d69159a)))


def f2_add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
def f2_sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
def f2_inv(a):
    n = pow(a[0] * a[0] + a[1] * a[1], P - 2, P)
    return (a[0] * n % P, -a[1] * n % P)


TWIST_B = f2_mul((3, 0), f2_inv((3, 1)))          # 3 / (i + 3)


def g1_add(a, b):
    if a is None: return b
    if b is None: return a
    x1, y1 = a; x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P == 0: return None
        lam = 3 * x1 * x1 * pow(2 * y1, P - 2, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, P - 2, P) % P
    x3 = (lam * lam - x1 - x2) % P
    return (x3, (lam * (x1 - x3) - y1) % P)


def g1_mul(k, pt=G1):
    acc = None
    for bit in bin(k % ORDER)[2:] if (pt is not None and k % ORDER) else "":
        acc = g1_add(acc, acc)
        if bit == "1": acc = g1_add(acc, pt)
    return acc


def g2_add(a, b):
    if a is None: return b
    if b is None: return a
    x1, y1 = a; x2, y2 = b
    if x1 == x2:
        if f2_add(y1, y2) == (0, 0): return None
        lam = f2_mul(f2_mul((3, 0), f2_mul(x1, x1)), f2_inv(f2_add(y1, y1)))
    else:
        lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_mul(lam, lam), x1), x2)
    return (x3, f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1))


def g2_mul(k, pt=G2):
    acc = None
    for bit in bin(k % ORDER)[2:] if (pt is not None and k % ORDER) else "":
        acc = g2_add(acc, acc)
        if bit == "1": acc = g2_add(acc, pt)
    return acc


def g1_marshal(pt) -> bytes:
    return bytes(64) if pt is None else pt[0].to_bytes(32, "big") + pt[1].to_bytes(32, "big")


def g1_unmarshal(b: bytes):
    if b == bytes(64): return None
    return (int.from_bytes(b[:32], "big"), int.from_bytes(b[32:], "big"))


def g2_marshal(pt) -> bytes:
    if pt is None: return bytes(128)
    (xr, xi), (yr, yi) = pt
    return b"".join(v.to_bytes(32, "big") for v in (xi, xr, yi, yr))


def g2_unmarshal(b: bytes):
    if b == bytes(128): return None
    v = [int.from_bytes(b[32 * i:32 * i + 32], "big") for i in range(4)]
    return ((v[1], v[0]), (v[3], v[2]))


def hash_to_g1(m: bytes):
    x = int.from_bytes(hashlib.sha256(m).digest(), "big") % P
    while True:
        t = (x ** 3 + B) % P
        y = pow(t, (P + 1) // 4, P)
        if y * y % P == t:
            return (x, y)
        x += 1


assert (G1[1] ** 2 - G1[0] ** 3 - B) % P == 0
assert f2_sub(f2_mul(G2[1], G2[1]), f2_add(f2_mul(f2_mul(G2[0], G2[0]), G2[0]), TWIST_B)) == (0, 0)
