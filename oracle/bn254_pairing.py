"""CPU ORACLE (test infrastructure, NOT product code) -- bn254 G2 and the optimal-ate pairing, following the
in-tree reference step by step (this curve's arithmetic IS in /root/reference, so the GT bytes are
source-pinned):
  tower        pairing/bn254/gfp2.go (x*i + y, i^2 = -1; MulXi: xi = i+9), gfp6.go (x t^2 + y t + z, t^3 = xi),
               gfp12.go (x w + y, w^2 = t)
  lines        lineFunctionAdd / lineFunctionDouble / mulLine      pairing/bn254/optate.go:5-114
  miller       NAF(6u+2) loop + Q1, -Q2 Frobenius steps            optate.go:117-207
  finalExponentiation  easy part + the y0..y6 addition chain       optate.go:212-261
  twistGen     twist.go:22-33 (Montgomery limbs, R = 2^256); GT MarshalBinary point.go:625-656 (384 B,
               x.x.x first); G2 MarshalBinary point.go:428-455 (x.imag||x.real||y.imag||y.real)
Element conventions here: Fp2 = (real, imag); Fp6 = (c0, c1, c2) for c0 + c1 t + c2 t^2; Fp12 = (c0, c1) for
c0 + c1 w -- i.e. the Go fields in reverse order.
Only tests/ may import this.  No reference fixture holds bn254 GT bytes (the reference compares against
gnark-crypto, not runnable here): parity is pinned by following the source, not by a KAT.
"""
from __future__ import annotations
from . import bn254 as _c
from .bn_pairing_generic import build


def _mont(*ws):
    return sum(w << (64 * i) for i, w in enumerate(ws)) * pow(1 << 256, -1, _c.P) % _c.P


# twistGen, pairing/bn254/twist.go:22-33 (Montgomery limbs, gfP2{x = imag, y = real}) as ((real, imag), (real, imag))
_G2 = ((_mont(0x8e83b5d102bc2026, 0xdceb1935497b0172, 0xfbb8264797811adf, 0x19573841af96503b),
        _mont(0xafb4737da84c6140, 0x6043dd5a5802d8c4, 0x09e950fc52a02f86, 0x14fef0833aea7b6b)),
       (_mont(0x619dfa9d886be9f6, 0xfe7fd297f59e9b78, 0xff9e1a62231b7dfe, 0x28fd7eebae9e4206),
        _mont(0x64095b56c71856ee, 0xdc57f922327d3cbb, 0x55f935be33351076, 0x0da4a0e693fd6482)))
# sixuPlus2NAF, optate.go:117-120 (a signed-digit form of 6u+2, least significant digit first; data)
_DIGITS = [0, 0, 0, 1, 0, 1, 0, -1, 0, 0, 1, -1, 0, 0, 1, 0, 0, 1, 1, 0, -1, 0, 0, 1, 0, -1, 0, 0, 0, 0, 1, 1,
           1, 0, 0, -1, 0, 0, 1, 0, 0, 0, 0, 0, -1, 0, 0, 1, 1, 0, 0, -1, 0, 0, 0, 1, 1, 0, -1, 0, 0, 1, 0, 1, 1]
_ns = build(_c.P, _c.ORDER, _c.U, (9, 1), _DIGITS, _G2)
globals().update({k: v for k, v in vars(_ns).items() if not k.startswith("_") or k in ("_line_add", "_line_double", "_mul_line")})
P, ORDER, U, G1 = _c.P, _c.ORDER, _c.U, _c.G1
