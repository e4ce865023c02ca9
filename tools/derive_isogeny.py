#!/usr/bin/env python3
"""Derive the 11-isogeny E' -> E of the BLS12381G1 SSWU suite and emit it as device constants.

Standalone (imports nothing from oracle/): finds the unique rational subgroup of order 11 of
E': y^2 = x^3 + A'x + B', applies Velu's formulas, normalises with u = 11 so that the codomain is exactly
y^2 = x^3 + 4, and writes the rational-map coefficients in Montgomery form (R = 2^384).
  python tools/derive_isogeny.py > kyber_b200/csrc/iso_g1_constants.cuh
The same derivation exists independently in oracle/h2c_bls12381.py; tests/test_abi_and_host.py checks
that both agree and the oracle copy is pinned by the reference's signature KATs.
"""
import sys

P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R_ORD = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
X_ABS = 0xd201000000010000
A = 0x144698a3b8e9433d693a02c96d4982b0ea985383ee66a8d8e8981aefd881ac98936f8da0e0f97f5cf428082d584c1d
B = 0x12e2908d11688030018b12e8753eee3b2016c1f0f24f4070a0b9c14fcef35ef55a23215a316ceaa5d1cc48e98e172be0


def add(p1, p2):
    if p1 is None: return p2
    if p2 is None: return p1
    x1, y1 = p1; x2, y2 = p2
    if x1 == x2:
        if (y1 + y2) % P == 0: return None
        lam = (3 * x1 * x1 + A) * pow(2 * y1, P - 2, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, P - 2, P) % P
    x3 = (lam * lam - x1 - x2) % P
    return (x3, (lam * (x1 - x3) - y1) % P)


def mul(k, pt):
    acc = None
    for bit in bin(k)[2:]:
        acc = add(acc, acc)
        if bit == "1": acc = add(acc, pt)
    return acc


def padd(a, b):
    n = max(len(a), len(b)); return [((a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0)) % P for i in range(n)]
def psub(a, b):
    n = max(len(a), len(b)); return [((a[i] if i < len(a) else 0) - (b[i] if i < len(b) else 0)) % P for i in range(n)]
def pmul(a, b):
    r = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        for j, y in enumerate(b): r[i + j] = (r[i + j] + x * y) % P
    return r
def pder(a): return [(i * a[i]) % P for i in range(1, len(a))]
def pmod(a, m):
    a = a[:]; dm = len(m) - 1; inv = pow(m[-1], P - 2, P)
    while len(a) - 1 >= dm:
        c = a[-1] * inv % P; sh = len(a) - 1 - dm
        for i in range(dm + 1): a[sh + i] = (a[sh + i] - c * m[i]) % P
        a.pop()
    return a


def derive():
    n = (X_ABS + 1) ** 2 // 3 * R_ORD
    x, t = 0, None
    while t is None:
        x += 1
        y = pow((x ** 3 + A * x + B) % P, (P + 1) // 4, P)
        if y * y % P != (x ** 3 + A * x + B) % P: continue
        q = mul(n // 121, (x, y))
        if q is None: continue
        q11 = mul(11, q)
        t = q if q11 is None else q11
    xs, q = [], t
    for _ in range(5):
        xs.append(q[0]); q = add(q, t)
    h = [1]
    for xq in xs: h = pmul(h, [(-xq) % P, 1])
    hp = pder(h)
    gv = pmod(pmul([2 * A % P, 0, 6], hp), h)
    gu = pmod(pmul([4 * B % P, 4 * A % P, 0, 4], hp), h)
    h2 = pmul(h, h)
    nx = padd(pmul([0, 1], h2), padd(pmul(gv, h), psub(pmul(gu, hp), pmul(pder(gu), h))))
    ny = psub(pmul(pder(nx), h), pmul([2], pmul(nx, hp)))
    while nx[-1] == 0: nx.pop()
    while ny[-1] == 0: ny.pop()
    i2, i3 = pow(121, P - 2, P), pow(1331, P - 2, P)
    return [c * i2 % P for c in nx], h2, [c * i3 % P for c in ny], pmul(h2, h)


def limbs(v): return ", ".join("0x%08xu" % ((v >> (32 * i)) & 0xFFFFFFFF) for i in range(12))


# ---- G2: 3-isogeny E2' -> E2 over Fp2 (elements are (c0, c1)) ----------------------------------------------
def f2add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
def f2sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
def f2neg(a): return (-a[0] % P, -a[1] % P)
def f2mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
def f2muls(a, k): return (a[0] * k % P, a[1] * k % P)
def f2inv(a):
    n = pow((a[0] * a[0] + a[1] * a[1]) % P, P - 2, P)
    return (a[0] * n % P, -a[1] * n % P)
Z2 = (0, 0)
A2, B2, ZZ2 = (0, 240), (1012, 1012), ((-2) % P, (-1) % P)


def derive_g2():
    """x0 = the Fp2-rational root of psi3 of E2', v = 2(3x0^2 + A), u = 4(x0^3 + A x0 + B)."""
    psi3 = [f2neg(f2mul(A2, A2)), f2muls(B2, 12), f2muls(A2, 6), Z2, (3, 0)]
    def pmul2(a, b):
        r = [Z2] * (len(a) + len(b) - 1)
        for i, x in enumerate(a):
            for j, y in enumerate(b): r[i + j] = f2add(r[i + j], f2mul(x, y))
        return r
    def pmod2(a, m):
        a = a[:]; dm = len(m) - 1; inv = f2inv(m[-1])
        while len(a) - 1 >= dm:
            c = f2mul(a[-1], inv); sh = len(a) - 1 - dm
            for i in range(dm + 1): a[sh + i] = f2sub(a[sh + i], f2mul(c, m[i]))
            a.pop()
        while len(a) > 1 and a[-1] == Z2: a.pop()
        return a or [Z2]
    r = [(1, 0)]
    for bit in bin(P * P)[2:]:
        r = pmod2(pmul2(r, r), psi3)
        if bit == "1": r = pmod2(pmul2(r, [Z2, (1, 0)]), psi3)
    a = r + [Z2] * (2 - len(r))
    a[1] = f2sub(a[1], (1, 0))
    while len(a) > 1 and a[-1] == Z2: a.pop()
    b = psi3
    while not (len(a) == 1 and a[0] == Z2): b, a = a, pmod2(b, a)
    assert len(b) == 2
    x0 = f2neg(f2mul(b[0], f2inv(b[1])))
    v = f2muls(f2add(f2muls(f2mul(x0, x0), 3), A2), 2)
    u = f2muls(f2add(f2add(f2mul(f2mul(x0, x0), x0), f2mul(A2, x0)), B2), 4)
    assert f2sub(A2, f2muls(v, 5)) == Z2
    assert f2mul(f2sub(B2, f2muls(f2add(u, f2mul(x0, v)), 7)), f2inv((4, 4))) == (729, 0)
    return x0, v, u


def main():
    xn, xd, yn, yd = derive()
    Rm = 1 << 384
    o = sys.stdout
    o.write("// GENERATED by tools/derive_isogeny.py -- do not edit.\n"
            "// 11-isogeny E' -> E (BLS12381G1_XMD:SHA-256_SSWU_RO_), coefficients low -> high, Montgomery form.\n"
            "#pragma once\n#include \"ptx.cuh\"\nnamespace b2k {\n#if defined(__CUDACC__)\n#define B2K_TABLE static __device__ const\n#else\n#define B2K_TABLE static const\n#endif\n")
    for name, poly in (("ISO_G1_XNUM", xn), ("ISO_G1_XDEN", xd), ("ISO_G1_YNUM", yn), ("ISO_G1_YDEN", yd)):
        o.write("B2K_TABLE uint32_t %s[%d][12] = {\n" % (name, len(poly)))
        for c in poly: o.write("  {%s},\n" % limbs(c * Rm % P))
        o.write("};\n")
    for name, v in (("SSWU_A", A), ("SSWU_B", B), ("SSWU_Z", 11), ("SSWU_NEG_B_OVER_A", (-B) * pow(A, P - 2, P) % P),
                    ("SSWU_B_OVER_ZA", B * pow(11 * A, P - 2, P) % P), ("TWO_POW_256", (1 << 256) % P)):
        o.write("B2K_TABLE uint32_t %s[12] = {%s};\n" % (name, limbs(v * Rm % P)))
    x0, v, u = derive_g2()
    o.write("// G2 (BLS12381G2_XMD:SHA-256_SSWU_RO_): Fp2 constants as {c0 limbs, c1 limbs}.  iso_map(x,y) with d = x - x0:\n"
            "//   X = (x d^2 + v d + u) / (3d)^2 ,  Y = -y (d^3 - v d - 2u) / (3d)^3   (Velu, normalised by u = -3)\n")
    for name, val in (("ISO_G2_X0", x0), ("ISO_G2_V", v), ("ISO_G2_U", u), ("ISO_G2_2U", f2muls(u, 2)), ("SSWU2_A", A2), ("SSWU2_B", B2),
                      ("SSWU2_Z", ZZ2), ("SSWU2_NEG_B_OVER_A", f2mul(f2neg(B2), f2inv(A2))),
                      ("SSWU2_B_OVER_ZA", f2mul(B2, f2inv(f2mul(ZZ2, A2))))):
        o.write("B2K_TABLE uint32_t %s[2][12] = {{%s}, {%s}};\n" % (name, limbs(val[0] * Rm % P), limbs(val[1] * Rm % P)))
    o.write("}  // namespace b2k\n")


if __name__ == "__main__":
    main()
