#!/usr/bin/env python3
"""Generate the program tables of the WARP-COOPERATIVE pairings and GT exponentiation (kyber_b200/csrc/coop_program_*.inc).

Why: the batch kernels run one pairing per thread, so a single Suite.Pair / ValidatePairing (kilic/suite.go:57-75; pairing/bn254/suite.go:
133-144; pairing/bn256/suite.go:99-109) costs one thread's latency (20-26 ms).  A pairing is ~20 000 Fp products with a dependency depth of
only ~800, so one WARP can run one pairing with its 32 lanes executing independent Fp operations in lock step.  The schedule is static (the
loop bits of |x| / the digits of 6u + 2 are public), so it is compiled HERE, once per curve (BLS12-381, bn254, bn256): the tower / Miller /
final-exponentiation formulas below run over a symbolic field that records every Fp operation (exact zeros and ones propagate, so sparse
operands cost nothing), a list scheduler packs the operations into rounds of <= 32 (one per lane; products and additions never share a
round; only operations within SLACK of the critical path are eligible, which keeps the number of live values low), a linear-scan allocator
maps the values to slots of shared memory, and the device side (coop_core.cuh, coop_pairing.cuh) is a 50-line interpreter:
    for every round: lane l decodes word [round][l] -> (op, dst, a, b), loads its operands from shared memory, computes, stores; __syncwarp().
Per curve: the 1-pair program (Pair), the 2-pair program (ValidatePairing: shared squarings, one final exponentiation) and the Fp12
square / product programs that GT.Mul loops over (coop_program_<curve>_gt.inc).
The SAME formulas run over plain integers and must reproduce the oracle's values; the ENCODED programs are then interpreted numerically
(reads of a round before its writes, exactly like the lanes) and must reproduce them again -- `--check`, run by tests/test_coop_program.py,
which also runs the device interpreter's per-lane core under host emulation over the tables.
Generation needs only public curve parameters (kyber_b200/build.py runs it; the tables are git-ignored); only --check imports the oracle.

Usage:  python tools/gen_coop_pairing.py            # rewrite kyber_b200/csrc/coop_program_*.inc
        python tools/gen_coop_pairing.py --check    # validate formulas + schedule + encoding against the oracle; exit 1 on mismatch or stale files
"""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# Generation needs nothing but the curve's public parameters (kyber_b200/build.py runs it: the product build must not touch oracle/);
# only --check imports the oracle, as the checker.
BLS_P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
X_ABS = 0xd201000000010000                 # BLS12-381: the curve parameter is x = -X_ABS


def _f2_pow_int(a, e):                                             # (a0 + a1 u)^e over the integers mod P, u^2 = -1
    r = (1, 0)
    while e:
        if e & 1:
            r = ((r[0] * a[0] - r[1] * a[1]) % P, (r[0] * a[1] + r[1] * a[0]) % P)
        a = ((a[0] * a[0] - a[1] * a[1]) % P, 2 * a[0] * a[1] % P)
        e >>= 1
    return r


def _bn_p(u): return 36 * u ** 4 + 36 * u ** 3 + 24 * u ** 2 + 6 * u + 1


# public parameters of the three pairing curves: modulus, 32-bit limbs of the device field, xi (Fp6 = Fp2[v]/(v^3 - xi)) and, for the
# Barreto-Naehrig curves, u and the signed digits of 6u + 2 (data of pairing/bn254/optate.go:117-120, pairing/bn256/optate.go:117-122)
CURVES = {
    "BLS": {"P": BLS_P, "limbs": 12, "XI": (1, 1)},
    "BN254": {"P": _bn_p(4965661367192848881), "limbs": 8, "XI": (9, 1), "U": 4965661367192848881,
              "NAF": [0, 0, 0, 1, 0, 1, 0, -1, 0, 0, 1, -1, 0, 0, 1, 0, 0, 1, 1, 0, -1, 0, 0, 1, 0, -1, 0, 0, 0, 0, 1, 1,
                      1, 0, 0, -1, 0, 0, 1, 0, 0, 0, 0, 0, -1, 0, 0, 1, 1, 0, 0, -1, 0, 0, 0, 1, 1, 0, -1, 0, 0, 1, 0, 1, 1]},
    "BN256": {"P": _bn_p(6518589491078791937), "limbs": 10, "XI": (3, 1), "U": 6518589491078791937,
              "NAF": [0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, -1, 0, 1, 0, 1, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, -1, 0,
                      1, 0, 0, 0, 1, 0, -1, 0, 0, 0, -1, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, -1, 0, -1, 0, 0, 0, 0, 1, 0, 0, 0, 1]},
}
P = BLS_P                                  # the modulus of the curve being compiled (set_curve)
XI = (1, 1)
_GAMMA = {}
CUR = "BLS"


def set_curve(name):
    global P, XI, _GAMMA, CUR
    c = CURVES[name]
    CUR, P, XI = name, c["P"], c["XI"]
    _GAMMA = {j: [_f2_pow_int(XI, k * (P ** j - 1) // 6) for k in range(6)] for j in (1, 2, 3)}   # xi^(k (p^j - 1)/6)

# ---- operations -----------------------------------------------------------------------------------------------------------------
NOP, MUL, SQR, ADD, SUB, NEG, DBL, MULC, INV = range(9)
LONG = {MUL, SQR, MULC, INV}
COST = {MUL: 400, SQR: 400, MULC: 400, INV: 40000, ADD: 60, SUB: 60, NEG: 40, DBL: 60}
SLACK = 4000                                # scheduling window below the critical path, in COST units (compile_program)


class _Zero:
    def __repr__(self): return "ZERO"


class _One:
    def __repr__(self): return "ONE"


ZERO, ONE = _Zero(), _One()


class Backend:
    """Fp arithmetic with exact-zero / exact-one propagation (sparse operands cost nothing)."""

    def mat(self, a):                      # materialise the sentinel ONE
        return self.one() if a is ONE else a

    def mul(self, a, b):
        if a is ZERO or b is ZERO: return ZERO
        if a is ONE: return b
        if b is ONE: return a
        return self._sqr(a) if a is b else self._mul(a, b)

    def sqr(self, a):
        if a is ZERO or a is ONE: return a
        return self._sqr(a)

    def add(self, a, b):
        if a is ZERO: return b
        if b is ZERO: return a
        return self._dbl(self.mat(a)) if a is b else self._add(self.mat(a), self.mat(b))

    def sub(self, a, b):
        if b is ZERO: return a
        if a is ZERO: return self._neg(self.mat(b))
        if a is b: return ZERO
        return self._sub(self.mat(a), self.mat(b))

    def neg(self, a):
        return ZERO if a is ZERO else self._neg(self.mat(a))

    def dbl(self, a):
        return ZERO if a is ZERO else self._dbl(self.mat(a))

    def mulc(self, a, k):
        k %= P
        if a is ZERO or k == 0: return ZERO
        if k == 1: return a
        return self._mulc(self.mat(a), k)

    def inv(self, a):
        assert a is not ZERO
        return ONE if a is ONE else self._inv(a)


class Num(Backend):
    def one(self): return 1
    def _mul(self, a, b): return a * b % P
    def _sqr(self, a): return a * a % P
    def _add(self, a, b): return (a + b) % P
    def _sub(self, a, b): return (a - b) % P
    def _neg(self, a): return -a % P
    def _dbl(self, a): return 2 * a % P
    def _mulc(self, a, k): return a * k % P
    def _inv(self, a): return pow(a, P - 2, P)


class Sym(Backend):
    """records tasks (op, a, b, dst) over value ids; inputs, the constant ONE and the constant ZERO (id n_inputs + 1) are pinned ids"""

    def __init__(self, n_inputs):
        self.tasks = []
        self.nv = n_inputs + 2
        self.one_id = n_inputs
        self.consts = []

    def one(self): return self.one_id

    def _new(self, op, a, b=0):
        d = self.nv
        self.nv += 1
        self.tasks.append((op, a, b, d))
        return d

    def _mul(self, a, b): return self._new(MUL, a, b)
    def _sqr(self, a): return self._new(SQR, a)
    def _add(self, a, b): return self._new(ADD, a, b)
    def _sub(self, a, b): return self._new(SUB, a, b)
    def _neg(self, a): return self._new(NEG, a)
    def _dbl(self, a): return self._new(DBL, a)
    def _inv(self, a): return self._new(INV, a)

    def _mulc(self, a, k):
        if k not in self.consts:
            self.consts.append(k)
        return self._new(MULC, a, self.consts.index(k))


# ---- tower over a backend B: Fp2 = (c0, c1), Fp6 = (c0, c1, c2) over v^3 = xi = 1 + u, Fp12 = (c0, c1) over w^2 = v ---------------------
F2Z = (ZERO, ZERO)


def f2_is_zero(a): return a[0] is ZERO and a[1] is ZERO
def f2_add(B, a, b): return (B.add(a[0], b[0]), B.add(a[1], b[1]))
def f2_sub(B, a, b): return (B.sub(a[0], b[0]), B.sub(a[1], b[1]))
def f2_neg(B, a): return (B.neg(a[0]), B.neg(a[1]))
def f2_dbl(B, a): return (B.dbl(a[0]), B.dbl(a[1]))
def f2_conj(B, a): return (a[0], B.neg(a[1]))
def smul(B, a, k):                                                 # k a for a small positive integer k, by doublings and additions
    r, t = ZERO, a
    while k:
        if k & 1: r = B.add(r, t)
        k >>= 1
        if k: t = B.dbl(t)
    return r


def f2_smul(B, a, k): return (smul(B, a[0], k), smul(B, a[1], k))


def f2_mul_xi(B, a):                                               # (a0 + a1 u)(x0 + u): xi = x0 + u with x0 = 1 (BLS), 9 (bn254), 3 (bn256)
    assert XI[1] == 1
    return (B.sub(smul(B, a[0], XI[0]), a[1]), B.add(smul(B, a[1], XI[0]), a[0]))
def f2_mul_fp(B, a, k): return (B.mul(a[0], k), B.mul(a[1], k))


def f2_mul(B, a, b):
    if f2_is_zero(a) or f2_is_zero(b): return F2Z
    if any(x is ZERO for x in (a[0], a[1], b[0], b[1])):           # sparse: schoolbook, the zero terms vanish
        return (B.sub(B.mul(a[0], b[0]), B.mul(a[1], b[1])), B.add(B.mul(a[0], b[1]), B.mul(a[1], b[0])))
    t0, t1 = B.mul(a[0], b[0]), B.mul(a[1], b[1])
    s = B.mul(B.add(a[0], a[1]), B.add(b[0], b[1]))
    return (B.sub(t0, t1), B.sub(B.sub(s, t0), t1))


def f2_sqr(B, a):
    if a[1] is ZERO: return (B.sqr(a[0]), ZERO)
    if a[0] is ZERO: return (B.neg(B.sqr(a[1])), ZERO)
    m = B.mul(a[0], a[1])
    return (B.mul(B.add(a[0], a[1]), B.sub(a[0], a[1])), B.dbl(m))


def f2_mul_const(B, a, g):                                         # g = (g0, g1) integers
    if g[1] % P == 0: return (B.mulc(a[0], g[0]), B.mulc(a[1], g[0]))
    t0, t1 = B.mulc(a[0], g[0]), B.mulc(a[1], g[1])
    s = B.mulc(B.add(a[0], a[1]), (g[0] + g[1]) % P)
    return (B.sub(t0, t1), B.sub(B.sub(s, t0), t1))


def f2_inv(B, a):
    n = B.inv(B.add(B.sqr(a[0]), B.sqr(a[1])))
    return (B.mul(a[0], n), B.neg(B.mul(a[1], n)))


F6Z = (F2Z, F2Z, F2Z)


def f6_add(B, a, b): return tuple(f2_add(B, x, y) for x, y in zip(a, b))
def f6_sub(B, a, b): return tuple(f2_sub(B, x, y) for x, y in zip(a, b))
def f6_neg(B, a): return tuple(f2_neg(B, x) for x in a)
def f6_mul_v(B, a): return (f2_mul_xi(B, a[2]), a[0], a[1])


def f6_mul(B, a, b):
    za, zb = [f2_is_zero(x) for x in a], [f2_is_zero(x) for x in b]
    if sum(za) >= 2 or sum(zb) >= 2:                               # very sparse: schoolbook
        def m(i, j): return f2_mul(B, a[i], b[j])
        c0 = f2_add(B, m(0, 0), f2_mul_xi(B, f2_add(B, m(1, 2), m(2, 1))))
        c1 = f2_add(B, f2_add(B, m(0, 1), m(1, 0)), f2_mul_xi(B, m(2, 2)))
        c2 = f2_add(B, f2_add(B, m(0, 2), m(1, 1)), m(2, 0))
        return (c0, c1, c2)
    v0, v1, v2 = f2_mul(B, a[0], b[0]), f2_mul(B, a[1], b[1]), f2_mul(B, a[2], b[2])
    t0 = f2_sub(B, f2_sub(B, f2_mul(B, f2_add(B, a[1], a[2]), f2_add(B, b[1], b[2])), v1), v2)
    t1 = f2_sub(B, f2_sub(B, f2_mul(B, f2_add(B, a[0], a[1]), f2_add(B, b[0], b[1])), v0), v1)
    t2 = f2_sub(B, f2_sub(B, f2_mul(B, f2_add(B, a[0], a[2]), f2_add(B, b[0], b[2])), v0), v2)
    return (f2_add(B, v0, f2_mul_xi(B, t0)), f2_add(B, t1, f2_mul_xi(B, v2)), f2_add(B, t2, v1))


def f6_inv(B, a):
    c0 = f2_sub(B, f2_sqr(B, a[0]), f2_mul_xi(B, f2_mul(B, a[1], a[2])))
    c1 = f2_sub(B, f2_mul_xi(B, f2_sqr(B, a[2])), f2_mul(B, a[0], a[1]))
    c2 = f2_sub(B, f2_sqr(B, a[1]), f2_mul(B, a[0], a[2]))
    t = f2_add(B, f2_mul_xi(B, f2_add(B, f2_mul(B, a[2], c1), f2_mul(B, a[1], c2))), f2_mul(B, a[0], c0))
    ti = f2_inv(B, t)
    return (f2_mul(B, c0, ti), f2_mul(B, c1, ti), f2_mul(B, c2, ti))


F12_ONE = (((ONE, ZERO), F2Z, F2Z), F6Z)


def f12_mul(B, a, b):
    t0, t1 = f6_mul(B, a[0], b[0]), f6_mul(B, a[1], b[1])
    c1 = f6_sub(B, f6_sub(B, f6_mul(B, f6_add(B, a[0], a[1]), f6_add(B, b[0], b[1])), t0), t1)
    return (f6_add(B, t0, f6_mul_v(B, t1)), c1)


def f12_sqr(B, a):
    ab = f6_mul(B, a[0], a[1])
    t = f6_mul(B, f6_add(B, a[0], a[1]), f6_add(B, a[0], f6_mul_v(B, a[1])))
    return (f6_sub(B, f6_sub(B, t, ab), f6_mul_v(B, ab)), f6_add(B, ab, ab))


def f12_conj(B, a): return (a[0], f6_neg(B, a[1]))


def f12_inv(B, a):
    t = f6_sub(B, f6_mul(B, a[0], a[0]), f6_mul_v(B, f6_mul(B, a[1], a[1])))
    ti = f6_inv(B, t)
    return (f6_mul(B, a[0], ti), f6_neg(B, f6_mul(B, a[1], ti)))


# w-power slots: w^0 c0.c0, w^1 c1.c0, w^2 c0.c1, w^3 c1.c1, w^4 c0.c2, w^5 c1.c2
def _to_w(a): return [a[0][0], a[1][0], a[0][1], a[1][1], a[0][2], a[1][2]]
def _from_w(c): return ((c[0], c[2], c[4]), (c[1], c[3], c[5]))


def f12_frobenius(B, a, j):
    c = _to_w(a)
    out = []
    for k in range(6):
        x = f2_conj(B, c[k]) if j % 2 else c[k]
        out.append(f2_mul_const(B, x, _GAMMA[j][k]))
    return _from_w(out)


def fp4_sqr(B, a, b):
    t0, t1 = f2_sqr(B, a), f2_sqr(B, b)
    c0 = f2_add(B, f2_mul_xi(B, t1), t0)
    c1 = f2_sub(B, f2_sub(B, f2_sqr(B, f2_add(B, a, b)), t0), t1)
    return c0, c1


def f12_cyclotomic_sqr(B, f):                                      # Granger-Scott (tower.cuh: fp12_cyclotomic_sqr)
    z0, z4, z3, z2, z1, z5 = f[0][0], f[0][1], f[0][2], f[1][0], f[1][1], f[1][2]

    def fix(t, z, plus):                                           # 3 t +- 2 z
        u = f2_add(B, t, z) if plus else f2_sub(B, t, z)
        return f2_add(B, f2_dbl(B, u), t)
    t0, t1 = fp4_sqr(B, z0, z1)
    t2, t3 = fp4_sqr(B, z2, z3)
    t4, t5 = fp4_sqr(B, z4, z5)
    t5 = f2_mul_xi(B, t5)
    return ((fix(t0, z0, False), fix(t2, z4, False), fix(t4, z3, False)),
            (fix(t5, z2, True), fix(t1, z1, True), fix(t3, z5, True)))


def f12_pow_x(B, a):                                               # a^x, x = -|x|, inside the cyclotomic subgroup
    acc = a
    for bit in bin(X_ABS)[3:]:
        acc = f12_cyclotomic_sqr(B, acc)
        if bit == "1":
            acc = f12_mul(B, acc, a)
    return f12_conj(B, acc)


def final_exponentiation(B, f):                                    # exponent 3 (p^12 - 1)/r  (pairing.cuh: final_exponentiation)
    m = f12_mul(B, f12_conj(B, f), f12_inv(B, f))
    m = f12_mul(B, f12_frobenius(B, m, 2), m)
    b = f12_mul(B, f12_pow_x(B, m), f12_conj(B, m))
    a = f12_mul(B, f12_pow_x(B, b), f12_conj(B, b))
    c = f12_mul(B, f12_pow_x(B, a), f12_frobenius(B, a, 1))
    a = f12_pow_x(B, f12_pow_x(B, c))
    a = f12_mul(B, f12_mul(B, a, f12_frobenius(B, c, 2)), f12_conj(B, c))
    t = f12_mul(B, f12_cyclotomic_sqr(B, m), m)
    return f12_mul(B, a, t)


# ---- Miller loop (pairing.cuh: Jacobian twist point, inversion-free lines (l0, l2, 0) + (0, l3, 0) w) ----------------------------------
def double_step(B, T, Pt):
    X, Y, Z = T
    A, Bq, ZZ = f2_sqr(B, X), f2_sqr(B, Y), f2_sqr(B, Z)
    C = f2_sqr(B, Bq)
    D = f2_dbl(B, f2_sub(B, f2_sub(B, f2_sqr(B, f2_add(B, X, Bq)), A), C))
    E = f2_add(B, f2_dbl(B, A), A)
    l0 = f2_sub(B, f2_sub(B, f2_mul(B, E, X), Bq), Bq)
    l2 = f2_neg(B, f2_mul_fp(B, f2_mul(B, E, ZZ), Pt[0]))
    Z3 = f2_dbl(B, f2_mul(B, Y, Z))
    l3 = f2_mul_fp(B, f2_mul(B, Z3, ZZ), Pt[1])
    X3 = f2_sub(B, f2_sub(B, f2_sqr(B, E), D), D)
    C8 = f2_dbl(B, f2_dbl(B, f2_dbl(B, C)))
    Y3 = f2_sub(B, f2_mul(B, E, f2_sub(B, D, X3)), C8)
    return (X3, Y3, Z3), (l0, l2, l3)


def add_step(B, T, Q, Pt):
    X, Y, Z = T
    ZZ = f2_sqr(B, Z)
    U2 = f2_mul(B, Q[0], ZZ)
    S2 = f2_mul(B, f2_mul(B, Q[1], Z), ZZ)
    H, Rr = f2_sub(B, U2, X), f2_sub(B, S2, Y)
    HH = f2_sqr(B, H)
    HHH = f2_mul(B, H, HH)
    V = f2_mul(B, X, HH)
    Z3 = f2_mul(B, Z, H)
    l0 = f2_sub(B, f2_mul(B, Rr, Q[0]), f2_mul(B, Z3, Q[1]))
    l2 = f2_neg(B, f2_mul_fp(B, Rr, Pt[0]))
    l3 = f2_mul_fp(B, Z3, Pt[1])
    X3 = f2_sub(B, f2_sub(B, f2_sub(B, f2_sqr(B, Rr), HHH), V), V)
    Y3 = f2_sub(B, f2_mul(B, Rr, f2_sub(B, V, X3)), f2_mul(B, Y, HHH))
    return (X3, Y3, Z3), (l0, l2, l3)


def mul_line(B, f, l):
    return f12_mul(B, f, ((l[0], l[1], F2Z), (F2Z, l[2], F2Z)))


def miller_loop(B, pairs):                                         # pairs: [(P = (x, y), Q = ((x0, x1), (y0, y1)))]
    T = [(Q[0], Q[1], (ONE, ZERO)) for _, Q in pairs]
    f = F12_ONE
    for bit in bin(X_ABS)[3:]:
        f = f12_sqr(B, f)
        for i, (Pt, Q) in enumerate(pairs):
            T[i], l = double_step(B, T[i], Pt)
            f = mul_line(B, f, l)
        if bit == "1":
            for i, (Pt, Q) in enumerate(pairs):
                T[i], l = add_step(B, T[i], Q, Pt)
                f = mul_line(B, f, l)
    return f12_conj(B, f)


# ---- Barreto-Naehrig curves (bn254, bn256): the reference's own algorithm, pairing/bn254/optate.go (bn256 twin) -----------------------------
# twist point r = (x, y, z, t = z^2) Jacobian; line = (a t + b) w + c with a, b in Fp2 and c in Fp2 (sparse product: mul_line_bn)
def bn_line_add(B, r, p, q, r2):                                   # optate.go:5-54; p = affine twist point, q = affine G1 point
    rx, ry, rz, rt = r
    Bv = f2_mul(B, p[0], rt)
    D = f2_add(B, p[1], rz)
    D = f2_mul(B, f2_sub(B, f2_sub(B, f2_sqr(B, D), r2), rt), rt)
    H = f2_sub(B, Bv, rx)
    I = f2_sqr(B, H)
    E = f2_smul(B, I, 4)
    J = f2_mul(B, H, E)
    L1 = f2_sub(B, f2_sub(B, D, ry), ry)
    V = f2_mul(B, rx, E)
    ox = f2_sub(B, f2_sub(B, f2_sub(B, f2_sqr(B, L1), J), V), V)
    oz = f2_sub(B, f2_sub(B, f2_sqr(B, f2_add(B, rz, H)), rt), I)
    t = f2_mul(B, f2_sub(B, V, ox), L1)
    t2 = f2_dbl(B, f2_mul(B, ry, J))
    oy = f2_sub(B, t, t2)
    ot = f2_sqr(B, oz)
    t = f2_sub(B, f2_sub(B, f2_sqr(B, f2_add(B, p[1], oz)), r2), ot)
    t2 = f2_dbl(B, f2_mul(B, L1, p[0]))
    a = f2_sub(B, t2, t)
    c = f2_dbl(B, f2_mul_fp(B, oz, q[1]))
    b = f2_dbl(B, f2_mul_fp(B, f2_neg(B, L1), q[0]))
    return a, b, c, (ox, oy, oz, ot)


def bn_line_double(B, r, q):                                       # optate.go:56-94
    rx, ry, rz, rt = r
    A, Bq = f2_sqr(B, rx), f2_sqr(B, ry)
    C = f2_sqr(B, Bq)
    D = f2_dbl(B, f2_sub(B, f2_sub(B, f2_sqr(B, f2_add(B, rx, Bq)), A), C))
    E = f2_smul(B, A, 3)
    G = f2_sqr(B, E)
    ox = f2_sub(B, f2_sub(B, G, D), D)
    oz = f2_sub(B, f2_sub(B, f2_sqr(B, f2_add(B, ry, rz)), Bq), rt)
    oy = f2_sub(B, f2_mul(B, f2_sub(B, D, ox), E), f2_smul(B, C, 8))
    ot = f2_sqr(B, oz)
    t = f2_dbl(B, f2_mul(B, E, rt))
    b = f2_mul_fp(B, f2_neg(B, t), q[0])
    a = f2_sub(B, f2_sub(B, f2_sub(B, f2_sqr(B, f2_add(B, rx, E)), A), G), f2_smul(B, Bq, 4))
    c = f2_mul_fp(B, f2_dbl(B, f2_mul(B, oz, rt)), q[1])
    return a, b, c, (ox, oy, oz, ot)


def bn_mul_line(B, f, a, b, c):                                    # f * ((a t + b) w + c), optate.go:96-114
    return f12_mul(B, f, ((c, F2Z, F2Z), (b, a, F2Z)))


def bn_miller(B, pairs):                                           # pairs: [(P = (x, y), Q = ((x0, x1), (y0, y1)))]; shared squarings
    naf = CURVES[CUR]["NAF"]
    n = len(naf)
    R = [(Q[0], Q[1], (ONE, ZERO), (ONE, ZERO)) for _, Q in pairs]
    R2 = [f2_sqr(B, Q[1]) for _, Q in pairs]
    f = F12_ONE
    for i in range(n - 1, 0, -1):
        if i != n - 1:
            f = f12_sqr(B, f)
        for k, (Pt, Q) in enumerate(pairs):
            a, b, c, R[k] = bn_line_double(B, R[k], Pt)
            f = bn_mul_line(B, f, a, b, c)
        d = naf[i - 1]
        if d:
            for k, (Pt, Q) in enumerate(pairs):
                Qs = Q if d == 1 else (Q[0], f2_neg(B, Q[1]))
                a, b, c, R[k] = bn_line_add(B, R[k], Qs, Pt, R2[k])
                f = bn_mul_line(B, f, a, b, c)
    xi_p1_3, xi_p1_2 = _f2_pow_int(XI, (P - 1) // 3), _f2_pow_int(XI, (P - 1) // 2)
    xi_p2_3 = _f2_pow_int(XI, (P * P - 1) // 3)
    assert xi_p2_3[1] == 0
    for k, (Pt, Q) in enumerate(pairs):                            # the two Frobenius steps: Q1 = pi(Q), -Q2 = -pi^2(Q)
        q1 = (f2_mul_const(B, f2_conj(B, Q[0]), xi_p1_3), f2_mul_const(B, f2_conj(B, Q[1]), xi_p1_2))
        mq2 = (f2_mul_const(B, Q[0], xi_p2_3), Q[1])
        a, b, c, R[k] = bn_line_add(B, R[k], q1, Pt, f2_sqr(B, q1[1]))
        f = bn_mul_line(B, f, a, b, c)
        a, b, c, _ = bn_line_add(B, R[k], mq2, Pt, f2_sqr(B, mq2[1]))
        f = bn_mul_line(B, f, a, b, c)
    return f


def f12_pow_u(B, a, u):                                            # a^u inside the cyclotomic subgroup
    acc = a
    for bit in bin(u)[3:]:
        acc = f12_cyclotomic_sqr(B, acc)
        if bit == "1":
            acc = f12_mul(B, acc, a)
    return acc


def bn_final_exponentiation(B, f):                                 # optate.go:212-261
    U = CURVES[CUR]["U"]
    t1 = f12_mul(B, f12_conj(B, f), f12_inv(B, f))
    t1 = f12_mul(B, t1, f12_frobenius(B, t1, 2))
    fp, fp2, fp3 = f12_frobenius(B, t1, 1), f12_frobenius(B, t1, 2), f12_frobenius(B, t1, 3)
    fu = f12_pow_u(B, t1, U)
    fu2 = f12_pow_u(B, fu, U)
    fu3 = f12_pow_u(B, fu2, U)
    y3 = f12_conj(B, f12_frobenius(B, fu, 1))
    fu2p, fu3p = f12_frobenius(B, fu2, 1), f12_frobenius(B, fu3, 1)
    y2 = f12_frobenius(B, fu2, 2)
    y0 = f12_mul(B, f12_mul(B, fp, fp2), fp3)
    y1, y5 = f12_conj(B, t1), f12_conj(B, fu2)
    y4 = f12_conj(B, f12_mul(B, fu, fu2p))
    y6 = f12_conj(B, f12_mul(B, fu3, fu3p))
    t0 = f12_mul(B, f12_mul(B, f12_cyclotomic_sqr(B, y6), y4), y5)
    t1b = f12_mul(B, f12_mul(B, y3, y5), t0)
    t0 = f12_mul(B, t0, y2)
    t1b = f12_cyclotomic_sqr(B, f12_mul(B, f12_cyclotomic_sqr(B, t1b), t0))
    t0 = f12_mul(B, t1b, y1)
    t1b = f12_mul(B, t1b, y0)
    return f12_mul(B, f12_cyclotomic_sqr(B, t0), t1b)


def pairing_product(B, pairs):
    if CUR == "BLS":
        return final_exponentiation(B, miller_loop(B, pairs))
    return bn_final_exponentiation(B, bn_miller(B, pairs))


def flat12(f):                                                     # 12 Fp values in the order of BFp12's memory layout
    return [f[h][k][c] for h in range(2) for k in range(3) for c in range(2)]


# ---- compile: schedule into rounds, allocate slots, encode ---------------------------------------------------------------------------
def unflat12(v):                                                   # inverse of flat12 on a list of 12 values
    return tuple(tuple((v[6 * h + 2 * k], v[6 * h + 2 * k + 1]) for k in range(3)) for h in range(2))


def compile_program(curve, kind):
    """kind 1 / 2: the product of that many pairings incl. the final exponentiation (inputs: 6 values per pair);
    kind "SQR" / "MUL": the square of one / the product of two Fp12 elements (inputs: 12 / 24 values) -- the steps of GT.Mul"""
    set_curve(curve)
    npairs = kind
    if kind in ("SQR", "MUL"):
        n_in = 12 if kind == "SQR" else 24
        B = Sym(n_in)
        a = unflat12(list(range(12)))
        out = flat12(f12_sqr(B, a) if kind == "SQR" else f12_mul(B, a, unflat12(list(range(12, 24)))))
    else:
        n_in = 6 * npairs                                          # per pair: P.x, P.y, Q.x.c0, Q.x.c1, Q.y.c0, Q.y.c1
        B = Sym(n_in)
        pairs = [((6 * i, 6 * i + 1), ((6 * i + 2, 6 * i + 3), (6 * i + 4, 6 * i + 5))) for i in range(npairs)]
        out = flat12(pairing_product(B, pairs))
    assert all(isinstance(x, int) for x in out), "an output coefficient is a constant"
    tasks = B.tasks
    producer = {t[3]: i for i, t in enumerate(tasks)}
    pinned = set(range(n_in + 2))
    zero_id = n_in + 1

    def srcs(t):
        op, a, b, d = t
        return (a, b) if op in (MUL, ADD, SUB) else (a,)
    # dead-code elimination from the outputs
    live, stack = set(), [producer[v] for v in out if v in producer]
    while stack:
        i = stack.pop()
        if i in live: continue
        live.add(i)
        for v in srcs(tasks[i]):
            if v in producer: stack.append(producer[v])
    tasks = [t for i, t in enumerate(tasks) if i in live]
    producer = {t[3]: i for i, t in enumerate(tasks)}
    users = {}
    for i, t in enumerate(tasks):
        for v in set(srcs(t)):
            users.setdefault(v, []).append(i)
    # priority = longest path to an output
    prio = [0] * len(tasks)
    for i in range(len(tasks) - 1, -1, -1):
        best = 0
        for u in users.get(tasks[i][3], []):
            best = max(best, prio[u])
        prio[i] = best + COST[tasks[i][0]]
    ndeps = [sum(1 for v in set(srcs(t)) if v in producer) for t in tasks]
    ready_long = [i for i, t in enumerate(tasks) if ndeps[i] == 0 and t[0] in LONG]
    ready_short = [i for i, t in enumerate(tasks) if ndeps[i] == 0 and t[0] not in LONG]
    rounds, done = [], 0
    while done < len(tasks):
        # Only operations within SLACK of the longest remaining path are eligible: work that is ready early but needed late (the side
        # branches of the final exponentiation) would otherwise sit in slots for thousands of rounds (482 slots without the rule, 288 with
        # it, same number of rounds; the slot count decides how many warps share an SM).  Among the eligible ones short operations go
        # first: they feed the products and free slots early (products first was tried: as many rounds, 2.5 x the slots).
        pmax = max(prio[i] for i in ready_short + ready_long)
        elig_s = [i for i in ready_short if prio[i] >= pmax - SLACK]
        elig_l = [i for i in ready_long if prio[i] >= pmax - SLACK]
        pool_src, pool = (ready_short, elig_s) if elig_s else (ready_long, elig_l)
        pool.sort(key=lambda i: -prio[i])
        if pool_src is ready_long and tasks[pool[0]][0] == INV:      # an inversion runs alone (every other lane would idle anyway)
            take = [pool[0]]
        else:
            take = [i for i in pool if tasks[i][0] != INV][:32]
        for i in take:
            pool_src.remove(i)
        rounds.append(take)
        done += len(take)
        for i in take:
            for u in users.get(tasks[i][3], []):
                ndeps[u] -= 1
                if ndeps[u] == 0:
                    (ready_long if tasks[u][0] in LONG else ready_short).append(u)
    # slots: inputs + ONE pinned at 0..n_in; a value's slot is free again in the round AFTER its last read
    round_of = {}
    for r, take in enumerate(rounds):
        for i in take:
            round_of[i] = r
    last_use = {}
    for i, t in enumerate(tasks):
        for v in srcs(t):
            last_use[v] = max(last_use.get(v, -1), round_of[i])
    for v in out:
        last_use[v] = len(rounds) + 1
    slot = {v: v for v in pinned}
    free, nslots = [], n_in + 2
    expiring = {}
    for v, r in last_use.items():
        if v not in pinned:
            expiring.setdefault(r, []).append(v)
    prog = []
    for r, take in enumerate(rounds):
        for v in expiring.get(r - 1, []):
            free.append(slot[v])
        words = []
        for i in take:
            op, a, b, d = tasks[i]
            if free:
                s = free.pop()
            else:
                s = nslots
                nslots += 1
            slot[d] = s
            # the interpreter knows MUL, ADD, SUB, MULC, INV only: a^2 = a * a, 2 a = a + a, -a = 0 - a (pinned ZERO slot)
            sa = slot[a]
            if op == SQR: eop, ea, eb = MUL, sa, sa
            elif op == DBL: eop, ea, eb = ADD, sa, sa
            elif op == NEG: eop, ea, eb = SUB, zero_id, sa
            elif op == MULC: eop, ea, eb = MULC, sa, b
            elif op == INV: eop, ea, eb = INV, sa, 0
            else: eop, ea, eb = op, sa, slot[b]
            assert s < 512 and ea < 512 and eb < 512
            words.append((eop << 28) | (s << 18) | (ea << 9) | eb)
        prog.append(words + [0] * (32 - len(words)))
    return {"curve": curve, "npairs": npairs, "n_in": n_in, "one": n_in, "zero": zero_id, "rounds": prog, "nslots": nslots, "out": [slot[v] for v in out],
            "consts": B.consts, "ntasks": len(tasks),
            "nlong": sum(1 for take in rounds if tasks[take[0]][0] in LONG), "nmul": sum(1 for t in tasks if t[0] in LONG)}


def run_program(pg, inputs):
    """numeric interpreter of the ENCODED program: all reads of a round happen before its writes (like the lanes of a warp)"""
    S = [None] * pg["nslots"]
    for k, v in enumerate(inputs):
        S[k] = v % P
    S[pg["one"]] = 1
    S[pg["zero"]] = 0
    for words in pg["rounds"]:
        writes = []
        for w in words:
            op, d, a, b = w >> 28, (w >> 18) & 511, (w >> 9) & 511, w & 511
            if op == NOP: continue
            x = S[a]
            assert x is not None, "read of an unwritten slot"
            if op == MUL: y = x * S[b] % P
            elif op == SQR: y = x * x % P
            elif op == ADD: y = (x + S[b]) % P
            elif op == SUB: y = (x - S[b]) % P
            elif op == NEG: y = -x % P
            elif op == DBL: y = 2 * x % P
            elif op == MULC: y = x * pg["consts"][b] % P
            elif op == INV: y = pow(x, P - 2, P)
            writes.append((d, y))
        assert len({d for d, _ in writes}) == len(writes), "two lanes write one slot"
        for d, y in writes:
            S[d] = y
    return [S[s] for s in pg["out"]]


def flat_oracle(f):
    return [f[h][k][c] for h in range(2) for k in range(3) for c in range(2)]


def inputs_of(pairs_pts):
    v = []
    for p1, q2 in pairs_pts:
        v += [p1[0], p1[1], q2[0][0], q2[0][1], q2[1][0], q2[1][1]]
    return v


def _oracle(curve):
    """(random pair, a true 2-pair identity, the reference value of a product of pairings, the GT one) from the test oracle"""
    sys.path.insert(0, ROOT)
    import types
    if curve == "BLS":
        from oracle import bls12381 as o
        assert o.P == CURVES["BLS"]["P"] and o.X_ABS == X_ABS

        def product(pts):
            f = o.F12_ONE
            for p1, q2 in pts:
                f = o.f12_mul(f, o.miller_loop(p1, q2))
            return o.final_exponentiation_cubed(f)
        return types.SimpleNamespace(order=o.R, g1_mul=o.g1_mul, g2_mul=o.g2_mul, g1_neg=o.g1_neg, G2=o.G2, product=product, one=o.F12_ONE, f12_mul=o.f12_mul)
    if curve == "BN254":
        from oracle import bn254 as c, bn254_pairing as b
    else:
        from oracle import bn256 as c, bn256_pairing as b
    assert c.P == CURVES[curve]["P"] and c.U == CURVES[curve]["U"]

    def product(pts):
        f = b.F12_ONE
        for p1, q2 in pts:
            f = b.f12_mul(f, b.miller(q2, p1))
        return b.final_exponentiation(f)
    return types.SimpleNamespace(order=c.ORDER, g1_mul=c.g1_mul, g2_mul=b.g2_mul, g1_neg=lambda pt: (pt[0], -pt[1] % c.P), G2=b.G2,
                                 product=product, one=b.F12_ONE, f12_mul=b.f12_mul)


def check(programs):
    import random
    rng = random.Random(7)
    ok = True
    for (curve, npairs), pg in programs.items():
        set_curve(curve)
        o = _oracle(curve)
        if npairs in ("SQR", "MUL"):                               # Fp12 square / product on random elements against plain-integer tower formulas
            for trial in range(2):
                vals = [rng.randrange(P) for _ in range(pg["n_in"])]
                a = unflat12(vals[:12])
                num = flat12(f12_sqr(Num(), a) if npairs == "SQR" else f12_mul(Num(), a, unflat12(vals[12:])))
                want = flat_oracle(o.f12_mul(a, a if npairs == "SQR" else unflat12(vals[12:])))
                got = run_program(pg, vals)
                print(f"  {curve} Fp12 {npairs} program, trial {trial}: formulas {'ok' if num == want else 'WRONG'}, encoded program {'ok' if got == want else 'WRONG'}")
                ok = ok and num == want and got == want
            continue
        for trial in range(2):
            pts = [(o.g1_mul(rng.randrange(1, o.order)), o.g2_mul(rng.randrange(1, o.order))) for _ in range(npairs)]
            if trial == 1 and npairs == 2:                         # a true ValidatePairing instance: e(aG, bH) e(-abG, H) = 1
                a, b = rng.randrange(1, o.order), rng.randrange(1, o.order)
                pts = [(o.g1_mul(a), o.g2_mul(b)), (o.g1_neg(o.g1_mul(a * b % o.order)), o.G2)]
            want = flat_oracle(o.product(pts))
            num = flat12(pairing_product(Num(), [((p[0], p[1]), (q[0], q[1])) for p, q in pts]))
            num = [1 if x is ONE else 0 if x is ZERO else x for x in num]
            got = run_program(pg, inputs_of(pts))
            good = num == want and got == want
            if trial == 1 and npairs == 2:
                good = good and want == flat_oracle(o.one)
            print(f"  {curve} {npairs}-pair program, trial {trial}: formulas {'ok' if num == want else 'WRONG'}, encoded program {'ok' if got == want else 'WRONG'}")
            ok = ok and good
    return ok


def mont_words(k, limbs):
    v = k * (1 << (32 * limbs)) % P
    return [(v >> (32 * i)) & 0xffffffff for i in range(limbs)]


def emit(programs, curve, gt=False):
    """a table file of one curve: the pairing programs (kyber_b200/csrc/coop_program_<curve>.inc) or, gt=True, the Fp12 square / product
    programs of the GT group (coop_program_<curve>_gt.inc; its own file: another translation unit uses them)"""
    set_curve(curve)
    limbs = CURVES[curve]["limbs"]
    lines = [f"// GENERATED by tools/gen_coop_pairing.py -- do not edit.  Program tables of the warp-cooperative {curve} pairing (coop_pairing.cuh):",
             "// word = op << 28 | dst << 18 | a << 9 | b; 32 words (one per lane) per round; op 0 = idle lane.",
             "// ops: 1 MUL  3 ADD  4 SUB  7 MULC (b = constant index)  8 INV   (a^2, 2 a and -a are encoded as a * a, a + a and ZERO - a)",
             "#pragma once", "#include <stdint.h>", "namespace b2k { namespace coop {"]
    mine = {n: pg for (c, n), pg in programs.items() if c == curve and (n in ("SQR", "MUL")) == gt}
    cname = f"{curve}_GT_CONSTS" if gt else f"{curve}_CONSTS"
    consts = []
    for pg in mine.values():
        for k in pg["consts"]:
            if k not in consts:
                consts.append(k)
    lines.append(f"__device__ const uint32_t {cname}[{max(1, len(consts))}][{limbs}] = {{   // Montgomery form (R = 2^{32 * limbs}), little-endian 32-bit limbs")
    if not consts:
        lines.append("  {0},")
    for k in consts:
        lines.append("  {" + ", ".join("0x%08xu" % w for w in mont_words(k, limbs)) + "},")
    lines.append("};")
    for npairs, pg in mine.items():
        remap = {i: consts.index(k) for i, k in enumerate(pg["consts"])}
        tag = f"{curve}_{npairs}" if gt else f"{curve}_P{npairs}"
        lines.append(f"// {curve}, " + (f"Fp12 {npairs}" if gt else f"{npairs}-pair product + final exponentiation") + f": {pg['ntasks']} operations ({pg['nmul']} products) in {len(pg['rounds'])} rounds "
                     f"({pg['nlong']} of them product rounds), {pg['nslots']} slots")
        lines.append(f"__device__ const uint16_t {tag}_OUT[12] = {{" + ", ".join(str(x) for x in pg["out"]) + "};")
        lines.append(f"__device__ const uint32_t {tag}_PROG[{len(pg['rounds'])} * 32] = {{")
        for words in pg["rounds"]:
            fixed = []
            for w in words:
                if (w >> 28) == MULC:
                    w = (w & ~511) | remap[w & 511]
                fixed.append(w)
            lines.append("  " + ",".join("0x%xu" % w for w in fixed) + ",")
        lines.append("};")
        lines.append(f"struct {tag} {{")
        lines.append(f"  static constexpr int ROUNDS = {len(pg['rounds'])}, SLOTS = {pg['nslots']}, INPUTS = {pg['n_in']}, ONE = {pg['one']}, ZERO = {pg['zero']}, LIMBS = {limbs};")
        lines.append(f"  static __device__ __forceinline__ const uint32_t* prog() {{ return {tag}_PROG; }}")
        lines.append(f"  static __device__ __forceinline__ const uint16_t* out() {{ return {tag}_OUT; }}")
        lines.append(f"  static __device__ __forceinline__ const uint32_t* consts() {{ return &{cname}[0][0]; }}")
        lines.append("};")
    lines.append("} }  // namespace b2k::coop")
    return "\n".join(lines) + "\n"


def out_path(curve, gt=False): return os.path.join(ROOT, "kyber_b200", "csrc", f"coop_program_{curve.lower()}{'_gt' if gt else ''}.inc")


def main():
    programs = {(c, n): compile_program(c, n) for c in CURVES for n in (1, 2, "SQR", "MUL")}
    for (c, n), pg in programs.items():
        what = f"Fp12 {n}" if n in ("SQR", "MUL") else f"{n}-pair"
        print(f"{c} {what} program: {pg['ntasks']} operations, {pg['nmul']} products, {len(pg['rounds'])} rounds ({pg['nlong']} product rounds), "
              f"{pg['nslots']} slots, {len(pg['consts'])} constants")
    files = {out_path(c, gt): emit(programs, c, gt) for c in CURVES for gt in (False, True)}
    if "--check" in sys.argv:
        good = check(programs)
        same = all(os.path.exists(f) and open(f).read() == t for f, t in files.items())
        print("coop_program_*.inc are", "up to date" if same else "STALE (run tools/gen_coop_pairing.py)")
        sys.exit(0 if good and same else 1)
    for f, t in files.items():
        open(f, "w").write(t)
        print("wrote", f, len(t), "bytes, sha256", hashlib.sha256(t.encode()).hexdigest()[:16])


if __name__ == "__main__":
    main()
