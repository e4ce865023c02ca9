"""CPU ORACLE (test infrastructure, NOT product code) -- bn254 hash-to-G1 of the reference restated.

  pointG1.Hash / hashToPoint    pairing/bn254/point.go:208-218   hash_to_field (2 elements) -> map each -> add
  hashToField                   point.go:220-232                  expand_message_xmd (Keccak-256, 96 bytes), 48 bytes -> mod p
  mapToPoint                    point.go:234-285                  Shallue-van de Woestijne (RFC 9380 6.6.1), Z = 1
  expandMsgXmdKeccak256         point.go:289-331                  RFC 9380 5.3.1 with legacy Keccak-256 (rate 136)
  legendre / sgn0 / Sqrt        gfp.go:88-91,125-146              e^((p-1)/2); parity of the canonical value; e^((p+1)/4)
  constants c1..c4              constants.go:71-80                g(Z), -Z/2, sqrt(-g(Z)(3Z^2+4A)), -4g(Z)/(3Z^2+4A); derived here
No cofactor clearing (h = 1).  Default DSTs: suite.go:43-47.
Pinned by the reference's vectors (tests/golden/bn254_hash_vectors.json from point_test.go / test_vectors_test.go).
"""
from __future__ import annotations

from . import bn254 as o

P = o.P
B = 3
Z = 1
C1 = (Z ** 3 + B) % P                                   # g(Z) = 4
C2 = (-Z * pow(2, -1, P)) % P                           # -Z/2
C3 = pow((-C1 * (3 * Z * Z)) % P, (P + 1) // 4, P)      # sqrt(-g(Z) (3Z^2 + 4A)), A = 0: the (p+1)/4 power (even)
C4 = (-4 * C1 * pow(3 * Z * Z, -1, P)) % P              # -4 g(Z) / (3Z^2 + 4A)
assert C3 * C3 % P == (-12) % P and C3 % 2 == 0

# ---- legacy Keccak-256 (padding 0x01, not SHA3's 0x06) -------------------------------------------------------------
_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B, 0x0000000080000001,
       0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
       0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003, 0x8000000000008002, 0x8000000000000080,
       0x000000000000800A, 0x800000008000000A, 0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]   # [x][y]
_M64 = (1 << 64) - 1


def _rol(v, n):
    n %= 64
    return ((v << n) | (v >> (64 - n))) & _M64 if n else v


def _keccak_f(a):
    for rc in _RC:
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                b[y][(2 * x + 3 * y) % 5] = _rol(a[x][y], _ROT[x][y])
        a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        a[0][0] ^= rc
    return a


def keccak256(data: bytes) -> bytes:
    rate = 136
    msg = bytearray(data)
    msg.append(0x01)
    while len(msg) % rate:
        msg.append(0)
    msg[-1] |= 0x80
    a = [[0] * 5 for _ in range(5)]
    for off in range(0, len(msg), rate):
        for i in range(rate // 8):
            a[i % 5][i // 5] ^= int.from_bytes(msg[off + 8 * i:off + 8 * i + 8], "little")
        a = _keccak_f(a)
    return b"".join(a[i % 5][i // 5].to_bytes(8, "little") for i in range(4))


def expand_message_xmd_keccak(dst: bytes, msg: bytes, out_len: int) -> bytes:
    assert len(dst) <= 255
    dst_prime = dst + bytes([len(dst)])
    b0 = keccak256(bytes(136) + msg + out_len.to_bytes(2, "big") + b"\x00" + dst_prime)
    bi = keccak256(b0 + b"\x01" + dst_prime)
    out = bi
    ell = (out_len + 31) // 32
    for i in range(2, ell + 1):
        bi = keccak256(bytes(x ^ y for x, y in zip(b0, bi)) + bytes([i]) + dst_prime)
        out += bi
    return out[:out_len]


def hash_to_field(dst: bytes, msg: bytes):
    u = expand_message_xmd_keccak(dst, msg, 96)
    return int.from_bytes(u[:48], "big") % P, int.from_bytes(u[48:], "big") % P


def _g(x):
    return (x * x * x + B) % P


def _legendre(e):
    f = pow(e, (P - 1) // 2, P)
    return 0 if f == 0 else (1 if f == 1 else -1)


def map_to_point(u: int):
    tv1 = u * u % P * C1 % P
    tv2 = (1 + tv1) % P
    tv1 = (1 - tv1) % P
    tv3 = pow(tv1 * tv2 % P, P - 2, P)                    # inv0
    tv5 = u * tv1 % P * tv3 % P * C3 % P
    x1 = (C2 - tv5) % P
    x2 = (C2 + tv5) % P
    tv8 = tv2 * tv2 % P * tv3 % P
    x3 = (1 + C4 * (tv8 * tv8 % P)) % P
    if _legendre(_g(x1)) == 1:
        x = x1
    elif _legendre(_g(x2)) == 1:
        x = x2
    else:
        x = x3
    y = pow(_g(x), (P + 1) // 4, P)
    if (u & 1) != (y & 1):
        y = (-y) % P
    return (x, y)


def hash_to_g1(dst: bytes, msg: bytes):
    u0, u1 = hash_to_field(dst, msg)
    return o.g1_add(map_to_point(u0), map_to_point(u1))
