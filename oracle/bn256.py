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


# ---- HashG1: HKDF-SHA256 to the base field, then Shallue-van de Woestijne (pairing/bn256/hash.go:10-110) -----------------
# s = sqrt(-3): the root the reference uses, from its Montgomery limbs (pairing/bn256/constants.go:104-105, data; R = 2^256)
SVDW_S = (0x236e675956be783b | 0x053957e6f379ab64 << 64 | 0xe60789a768f4a5c4 << 128 | 0x04f8979dd8bad754 << 192) * pow(1 << 256, -1, P) % P
assert SVDW_S * SVDW_S % P == P - 3
SVDW_S_M1_HALF = (SVDW_S - 1) * pow(2, -1, P) % P


def hash_to_base(msg: bytes, dst: bytes = None) -> int:
    """gfp.go:46-67 hashToBase: 48 bytes of HKDF-SHA256(secret = msg, salt = dst, info = "H2C" 0 1), big-endian, mod p"""
    import hmac
    salt = dst if dst else bytes(32)                              # Go's hkdf: a nil salt is HashLen zero bytes
    prk = hmac.new(salt, msg, hashlib.sha256).digest()
    info = b"H2C\x00\x01"
    t1 = hmac.new(prk, info + b"\x01", hashlib.sha256).digest()
    t2 = hmac.new(prk, t1 + info + b"\x02", hashlib.sha256).digest()
    return int.from_bytes((t1 + t2)[:48], "big") % P


def _sign0(x: int) -> int:                                        # gfp.go:137-148
    return 1 if x >= (P - 1) // 2 else -1


def _legendre(x: int) -> int:                                     # gfp.go:150-162
    f = pow(x, (P - 1) // 2, P)
    return 0 if f == 0 else 2 * (f & 1) - 1


def map_to_curve(t: int):
    """hash.go:14-110, statement by statement (Fermat inverse: 0 -> 0)"""
    a = (1 + B + t * t) % P
    st = SVDW_S * t % P
    w0 = pow(st * a % P, P - 2, P)
    w = st * st % P * w0 % P
    e = _sign0(t)

    def finish(x):
        y = pow((x * x * x + B) % P, (P + 1) // 4, P)
        if e != _sign0(y):
            y = -y % P
        return (x, y)
    x1 = (SVDW_S_M1_HALF - t * w) % P
    if _legendre((x1 ** 3 + B) % P) == 1:
        return finish(x1)
    x2 = (-1 - x1) % P
    if _legendre((x2 ** 3 + B) % P) == 1:
        return finish(x2)
    x3 = (pow(a, 4, P) * w0 % P * w0 + 1) % P
    return finish(x3)


def hash_g1(msg: bytes, dst: bytes = None):
    return map_to_curve(hash_to_base(msg, dst))


assert (G1[1] ** 2 - G1[0] ** 3 - B) % P == 0
assert f2_sub(f2_mul(G2[1], G2[1]), f2_add(f2_mul(f2_mul(G2[0], G2[0]), G2[0]), TWIST_B)) == (0, 0)
