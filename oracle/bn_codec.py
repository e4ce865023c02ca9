"""CPU ORACLE (test infrastructure, NOT product code) -- which byte strings the reference's UnmarshalBinary accepts on
the BN curves.  Restates, rule by rule:
  bn254 G1  pairing/bn254/point.go:146-185: gfP.Unmarshal (gfp.go:101-119) rejects a coordinate >= p; (0,0) is the
            point at infinity; otherwise curvePoint.IsOnCurve (curve.go): y^2 = x^3 + 3
  bn254 G2  pairing/bn254/point.go:473-514: the same range rule on the four coordinates (x.imag, x.real, y.imag, y.real);
            all-zero = infinity; otherwise twistPoint.IsOnCurve (twist.go:50-66): on the twist AND Order * P = infinity
  bn256 G1  pairing/bn256/point.go:206-238: gfP.Unmarshal (gfp.go:115-122) has NO range check -- montEncode reduces
            mod p, so a coordinate equal to p reads as 0 --; on the curve y^2 = x^3 + 3
  bn256 G2  pairing/bn256/point.go:469-506: no range check, on the twist (twist.go:50-61); no order check
Only tests/ may import this."""
from __future__ import annotations
from . import bn254 as _c4
from . import bn254_pairing as _p4
from . import bn256 as _c6


def _ints(b: bytes, k: int):
    return [int.from_bytes(b[32 * i:32 * i + 32], "big") for i in range(k)]


def _mul_unreduced(add, k: int, pt):
    acc = None
    for bit in bin(k)[2:]:
        acc = add(acc, acc)
        if bit == "1":
            acc = add(acc, pt)
    return acc


def bn254_g1_ok(b: bytes) -> bool:
    x, y = _ints(b, 2)
    if x >= _c4.P or y >= _c4.P:
        return False
    if x == 0 and y == 0:
        return True
    return (y * y - x * x * x - 3) % _c4.P == 0


def bn254_g2_ok(b: bytes) -> bool:
    xi, xr, yi, yr = _ints(b, 4)
    if max(xi, xr, yi, yr) >= _c4.P:
        return False
    if xi == xr == yi == yr == 0:
        return True
    pt = ((xr, xi), (yr, yi))
    if not _p4.g2_is_on_curve(pt):
        return False
    return _mul_unreduced(_p4.g2_add, _c4.ORDER, pt) is None


def bn256_g1_ok(b: bytes) -> bool:
    x, y = (v % _c6.P for v in _ints(b, 2))
    if x == 0 and y == 0:
        return True
    return (y * y - x * x * x - 3) % _c6.P == 0


def bn256_g2_ok(b: bytes) -> bool:
    xi, xr, yi, yr = (v % _c6.P for v in _ints(b, 4))
    if xi == xr == yi == yr == 0:
        return True
    x, y = (xr, xi), (yr, yi)
    return _c6.f2_sub(_c6.f2_mul(y, y), _c6.f2_add(_c6.f2_mul(_c6.f2_mul(x, x), x), _c6.TWIST_B)) == (0, 0)
