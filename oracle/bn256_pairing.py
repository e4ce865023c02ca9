"""CPU ORACLE (test infrastructure, NOT product code) -- bn256 optimal-ate pairing: the twin of pairing/bn254
(pairing/bn256/optate.go is the same code with its own digit table :117-122; xi = i+3, gfp2.go:103-118;
u, p, Order constants.go:16-22; twistGen twist.go:22-33; GT MarshalBinary point.go, 384 B, x.x.x first).
Instantiates oracle/bn_pairing_generic.py; G1/G2 arithmetic and Hash are in oracle/bn256.py (pinned by the BDN
fixtures).  Only tests/ may import this."""
from __future__ import annotations
from . import bn256 as _c
from .bn_pairing_generic import build

# sixuPlus2NAF, pairing/bn256/optate.go:117-122 (data)
_DIGITS = [0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, -1, 0, 1, 0,
           1, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, -1, 0, 1, 0, 0, 0, 1, 0, -1,
           0, 0, 0, -1, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, -1, 0, -1, 0, 0, 0,
           0, 1, 0, 0, 0, 1]
_ns = build(_c.P, _c.ORDER, _c.U, (3, 1), _DIGITS, _c.G2)
globals().update({k: v for k, v in vars(_ns).items() if not k.startswith("_")})
P, ORDER, U, G1 = _c.P, _c.ORDER, _c.U, _c.G1
