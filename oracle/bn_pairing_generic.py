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
import types


def build(P, ORDER, U, XI, DIGITS, G2_GEN):
    """Instantiate the BN optimal-ate pairing of pairing/bn254 (and its twin pairing/bn256, which differs only in
    the prime, u, xi = i+3 and the digit table, pairing/bn256/optate.go:117-122) for one parameter set."""



    def f2_add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
    def f2_sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
    def f2_neg(a): return (-a[0] % P, -a[1] % P)
    def f2_conj(a): return (a[0], -a[1] % P)
    def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
    def f2_sqr(a): return f2_mul(a, a)
    def f2_muls(a, k): return (a[0] * k % P, a[1] * k % P)
    def f2_mul_xi(a): return f2_mul(a, XI)


    def f2_inv(a):
        n = pow(a[0] * a[0] + a[1] * a[1], P - 2, P)
        return (a[0] * n % P, -a[1] * n % P)


    def f2_pow(a, e):
        r = (1, 0)
        for bit in bin(e)[2:]:
            r = f2_sqr(r)
            if bit == "1":
                r = f2_mul(r, a)
        return r


    F6_ZERO = ((0, 0), (0, 0), (0, 0))
    F6_ONE = ((1, 0), (0, 0), (0, 0))


    def f6_add(a, b): return tuple(f2_add(x, y) for x, y in zip(a, b))
    def f6_sub(a, b): return tuple(f2_sub(x, y) for x, y in zip(a, b))
    def f6_neg(a): return tuple(f2_neg(x) for x in a)


    def f6_mul(a, b):
        a0, a1, a2 = a
        b0, b1, b2 = b
        c0 = f2_add(f2_mul(a0, b0), f2_mul_xi(f2_add(f2_mul(a1, b2), f2_mul(a2, b1))))
        c1 = f2_add(f2_add(f2_mul(a0, b1), f2_mul(a1, b0)), f2_mul_xi(f2_mul(a2, b2)))
        c2 = f2_add(f2_add(f2_mul(a0, b2), f2_mul(a1, b1)), f2_mul(a2, b0))
        return (c0, c1, c2)


    def f6_mul_tau(a): return (f2_mul_xi(a[2]), a[0], a[1])
    def f6_mul_f2(a, k): return tuple(f2_mul(x, k) for x in a)


    def f6_inv(a):
        a0, a1, a2 = a
        t0 = f2_sub(f2_sqr(a0), f2_mul_xi(f2_mul(a1, a2)))
        t1 = f2_sub(f2_mul_xi(f2_sqr(a2)), f2_mul(a0, a1))
        t2 = f2_sub(f2_sqr(a1), f2_mul(a0, a2))
        d = f2_add(f2_mul(a0, t0), f2_mul_xi(f2_add(f2_mul(a2, t1), f2_mul(a1, t2))))
        di = f2_inv(d)
        return (f2_mul(t0, di), f2_mul(t1, di), f2_mul(t2, di))


    F12_ONE = (F6_ONE, F6_ZERO)


    def f12_mul(a, b):
        t0 = f6_mul(a[0], b[0])
        t1 = f6_mul(a[1], b[1])
        c1 = f6_sub(f6_sub(f6_mul(f6_add(a[0], a[1]), f6_add(b[0], b[1])), t0), t1)
        return (f6_add(t0, f6_mul_tau(t1)), c1)


    def f12_sqr(a): return f12_mul(a, a)
    def f12_conj(a): return (a[0], f6_neg(a[1]))


    def f12_inv(a):
        d = f6_sub(f6_mul(a[0], a[0]), f6_mul_tau(f6_mul(a[1], a[1])))
        di = f6_inv(d)
        return (f6_mul(a[0], di), f6_neg(f6_mul(a[1], di)))


    def f12_pow(a, e):
        r = F12_ONE
        for bit in bin(e)[2:]:
            r = f12_sqr(r)
            if bit == "1":
                r = f12_mul(r, a)
        return r


    _G1C = [f2_pow(XI, k * (P - 1) // 6) for k in range(6)]


    def f12_frobenius(a, power=1):
        """a^(p^power) via conj + xi^(k(p-1)/6) on the w-power basis (gfp12.go:62-83 composes the same map)."""
        for _ in range(power):
            (a00, a01, a02), (a10, a11, a12) = a
            c = [a00, a10, a01, a11, a02, a12]
            c = [f2_mul(f2_conj(c[k]), _G1C[k]) for k in range(6)]
            a = ((c[0], c[2], c[4]), (c[1], c[3], c[5]))
        return a


    # ---- G2 ------------------------------------------------------------------------------------------------------
    TWIST_B = f2_mul((3, 0), f2_inv(XI))
    G2 = G2_GEN


    def g2_is_on_curve(pt):
        return pt is None or f2_sub(f2_sqr(pt[1]), f2_add(f2_mul(f2_sqr(pt[0]), pt[0]), TWIST_B)) == (0, 0)


    def g2_neg(pt): return None if pt is None else (pt[0], f2_neg(pt[1]))


    def g2_add(a, b):
        if a is None: return b
        if b is None: return a
        x1, y1 = a
        x2, y2 = b
        if x1 == x2:
            if f2_add(y1, y2) == (0, 0):
                return None
            lam = f2_mul(f2_muls(f2_sqr(x1), 3), f2_inv(f2_muls(y1, 2)))
        else:
            lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
        x3 = f2_sub(f2_sub(f2_sqr(lam), x1), x2)
        return (x3, f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1))


    def g2_mul(k, pt=G2):
        acc = None
        for bit in bin(k % ORDER)[2:] if (pt is not None and k % ORDER) else "":
            acc = g2_add(acc, acc)
            if bit == "1":
                acc = g2_add(acc, pt)
        return acc


    def g2_marshal(pt) -> bytes:
        if pt is None:
            return bytes(128)
        (xr, xi), (yr, yi) = pt
        return b"".join(v.to_bytes(32, "big") for v in (xi, xr, yi, yr))


    assert g2_is_on_curve(G2) and g2_mul(ORDER - 1) == g2_neg(G2)


    # ---- line functions, exactly as optate.go:5-94 (twist point r = (x, y, z, t = z^2) Jacobian) ------------------
    def _line_add(r, p, q, r2):
        rx, ry, rz, rt = r
        B = f2_mul(p[0], rt)
        D = f2_add(p[1], rz)
        D = f2_mul(f2_sub(f2_sub(f2_sqr(D), r2), rt), rt)
        H = f2_sub(B, rx)
        I = f2_sqr(H)
        E = f2_muls(I, 4)
        J = f2_mul(H, E)
        L1 = f2_sub(f2_sub(D, ry), ry)
        V = f2_mul(rx, E)
        ox = f2_sub(f2_sub(f2_sub(f2_sqr(L1), J), V), V)
        oz = f2_sub(f2_sub(f2_sqr(f2_add(rz, H)), rt), I)
        t = f2_mul(f2_sub(V, ox), L1)
        t2 = f2_muls(f2_mul(ry, J), 2)
        oy = f2_sub(t, t2)
        ot = f2_sqr(oz)
        t = f2_sub(f2_sub(f2_sqr(f2_add(p[1], oz)), r2), ot)
        t2 = f2_muls(f2_mul(L1, p[0]), 2)
        a = f2_sub(t2, t)
        c = f2_muls(f2_muls(oz, q[1]), 2)
        b = f2_muls(f2_muls(f2_neg(L1), q[0]), 2)
        return a, b, c, (ox, oy, oz, ot)


    def _line_double(r, q):
        rx, ry, rz, rt = r
        A = f2_sqr(rx)
        B = f2_sqr(ry)
        C = f2_sqr(B)
        D = f2_muls(f2_sub(f2_sub(f2_sqr(f2_add(rx, B)), A), C), 2)
        E = f2_muls(A, 3)
        G = f2_sqr(E)
        ox = f2_sub(f2_sub(G, D), D)
        oz = f2_sub(f2_sub(f2_sqr(f2_add(ry, rz)), B), rt)
        oy = f2_sub(f2_mul(f2_sub(D, ox), E), f2_muls(C, 8))
        ot = f2_sqr(oz)
        t = f2_muls(f2_mul(E, rt), 2)
        b = f2_muls(f2_neg(t), q[0])
        a = f2_sub(f2_sub(f2_sub(f2_sqr(f2_add(rx, E)), A), G), f2_muls(B, 4))
        c = f2_muls(f2_muls(f2_mul(oz, rt), 2), q[1])
        return a, b, c, (ox, oy, oz, ot)


    def _mul_line(f, a, b, c):
        """f * ((a t + b) w + c)   (optate.go:96-114)"""
        line = (((c[0], c[1]), (0, 0), (0, 0)), (b, a, (0, 0)))
        return f12_mul(f, line)


    SIX_U_PLUS_2_NAF = DIGITS
    assert sum(d << i for i, d in enumerate(SIX_U_PLUS_2_NAF)) == 6 * U + 2

    XI_P1_3 = f2_pow(XI, (P - 1) // 3)
    XI_P1_2 = f2_pow(XI, (P - 1) // 2)
    XI_P2_3 = f2_pow(XI, (P * P - 1) // 3)[0]


    def miller(q2, p1):
        ret = F12_ONE
        a_aff, b_aff = q2, p1
        r = (a_aff[0], a_aff[1], (1, 0), (1, 0))
        r2 = f2_sqr(a_aff[1])
        n = len(SIX_U_PLUS_2_NAF)
        for i in range(n - 1, 0, -1):
            a, b, c, new_r = _line_double(r, b_aff)
            if i != n - 1:
                ret = f12_sqr(ret)
            ret = _mul_line(ret, a, b, c)
            r = new_r
            d = SIX_U_PLUS_2_NAF[i - 1]
            if d == 1:
                a, b, c, new_r = _line_add(r, a_aff, b_aff, r2)
            elif d == -1:
                a, b, c, new_r = _line_add(r, (a_aff[0], f2_neg(a_aff[1])), b_aff, r2)
            else:
                continue
            ret = _mul_line(ret, a, b, c)
            r = new_r
        q1 = (f2_mul(f2_conj(a_aff[0]), XI_P1_3), f2_mul(f2_conj(a_aff[1]), XI_P1_2))
        minus_q2 = (f2_muls(a_aff[0], XI_P2_3), a_aff[1])
        a, b, c, new_r = _line_add(r, q1, b_aff, f2_sqr(q1[1]))
        ret = _mul_line(ret, a, b, c)
        r = new_r
        a, b, c, _ = _line_add(r, minus_q2, b_aff, f2_sqr(minus_q2[1]))
        return _mul_line(ret, a, b, c)


    def final_exponentiation(f):
        t1 = f12_mul(f12_conj(f), f12_inv(f))
        t1 = f12_mul(t1, f12_frobenius(t1, 2))
        fp, fp2 = f12_frobenius(t1, 1), f12_frobenius(t1, 2)
        fp3 = f12_frobenius(fp2, 1)
        fu = f12_pow(t1, U); fu2 = f12_pow(fu, U); fu3 = f12_pow(fu2, U)
        y3 = f12_frobenius(fu, 1); fu2p = f12_frobenius(fu2, 1); fu3p = f12_frobenius(fu3, 1)
        y2 = f12_frobenius(fu2, 2)
        y0 = f12_mul(f12_mul(fp, fp2), fp3)
        y1 = f12_conj(t1); y5 = f12_conj(fu2); y3 = f12_conj(y3)
        y4 = f12_conj(f12_mul(fu, fu2p))
        y6 = f12_conj(f12_mul(fu3, fu3p))
        t0 = f12_mul(f12_mul(f12_sqr(y6), y4), y5)
        t1 = f12_mul(f12_mul(y3, y5), t0)
        t0 = f12_mul(t0, y2)
        t1 = f12_sqr(f12_mul(f12_sqr(t1), t0))
        t0 = f12_mul(t1, y1)
        t1 = f12_mul(t1, y0)
        return f12_mul(f12_sqr(t0), t1)


    def pairing(p1, q2):
        """Suite.Pair(p1, p2) = optimalAte(p2.g, p1.g)  (pairing/bn254/suite.go:133-136, optate.go:263-271)"""
        if p1 is None or q2 is None:
            return F12_ONE
        return final_exponentiation(miller(q2, p1))


    def gt_to_bytes(f) -> bytes:
        """point.go:625-656: x.x.x, x.x.y, x.y.x, ... , y.z.y with gfP12 = x w + y, gfP6 = x t^2 + y t + z, gfP2 = x i + y"""
        out = b""
        for c6 in (f[1], f[0]):
            for c2 in (c6[2], c6[1], c6[0]):
                out += c2[1].to_bytes(32, "big") + c2[0].to_bytes(32, "big")
        return out


    assert final_exponentiation(F12_ONE) == F12_ONE

    return types.SimpleNamespace(**{k: v for k, v in locals().items() if not k.startswith("__")})
