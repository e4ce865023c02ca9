"""CPU ORACLE (test infrastructure, NOT product code) -- bn254 G1 and scalars in Python big integers.

Restates the in-tree reference (this curve's arithmetic IS in /root/reference):
  constants      pairing/bn254/constants.go:24-28 (Order, p), curve.go:17-22 (generator (1,2)), b = 3
  curvePoint.Add/Double/Mul   pairing/bn254/curve.go:76-218 (result is the affine sum regardless of the
                              Jacobian/GLV route the Go code takes)
  MarshalBinary  pairing/bn254/point.go:113-132: x||y, 32-byte big-endian each, infinity = 64 zero bytes
  UnmarshalBinary point.go:146-185: rejects coordinates >= p (gfp.go:101-119) and off-curve points
Only tests/ and bench.py's cpu_baseline leg may import this.
"""
from __future__ import annotations

U = 4965661367192848881
P = 36 * U ** 4 + 36 * U ** 3 + 24 * U ** 2 + 6 * U + 1
ORDER = 36 * U ** 4 + 36 * U ** 3 + 18 * U ** 2 + 6 * U + 1
assert P == 21888242871839275222246405745257275088696311157297823662689037894645226208583
assert ORDER == 21888242871839275222246405745257275088548364400416034343698204186575808495617
B = 3
G1 = (1, 2)


def g1_is_on_curve(pt) -> bool:
    return pt is None or (pt[1] * pt[1] - pt[0] ** 3 - B) % P == 0


def g1_neg(pt):
    return None if pt is None else (pt[0], -pt[1] % P)


def g1_add(a, b):
    if a is None: return b
    if b is None: return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, P - 2, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, P - 2, P) % P
    x3 = (lam * lam - x1 - x2) % P
    return (x3, (lam * (x1 - x3) - y1) % P)


def g1_mul(k: int, pt=G1):
    acc = None
    if pt is None:
        return None
    for bit in bin(k % ORDER)[2:] if k % ORDER else "":
        acc = g1_add(acc, acc)
        if bit == "1":
            acc = g1_add(acc, pt)
    return acc


def g1_marshal(pt) -> bytes:
    if pt is None:
        return bytes(64)
    return pt[0].to_bytes(32, "big") + pt[1].to_bytes(32, "big")


def g1_unmarshal(b: bytes):
    if len(b) != 64:
        raise ValueError("wrong length")
    if b == bytes(64):
        return None
    x, y = int.from_bytes(b[:32], "big"), int.from_bytes(b[32:], "big")
    if x >= P or y >= P:
        raise ValueError("coordinate not below modulus")
    if not g1_is_on_curve((x, y)):
        raise ValueError("not on curve")
    return (x, y)


assert g1_is_on_curve(G1) and g1_mul(ORDER - 1) == g1_neg(G1)
