"""Exact integer MODEL of a 381-bit Montgomery product carried by the FP64 pipe ("next" item 1).

Not product code and not measured: a specification that the next round's CUDA kernel must reproduce, checked on the CPU
against plain big-int arithmetic (tests/test_dfma_model.py).  It fixes the radix, the constants and the headroom of the
scheme whose issue-rate budget DESIGN.md section 4 states (the FP64 pipe of a B200 SM issues one DFMA per 2 cycles per
sub-partition, next to one IMAD.WIDE per 4 cycles on the FMA-heavy pipe: profiles/r01_pipe_probes.txt).

Representation: 8 limbs of 52 bits held as doubles (exact integers < 2^52), Montgomery radix R = 2^416.
Partial product of two limbs a, b < 2^52 with round-toward-zero fused multiply-adds (`fma.rz.f64`):
    hi  = fma_rz(a, b, 2^104)                 in [2^104, 2^105): ulp 2^52  ->  hi  = 2^104 + floor(ab / 2^52) 2^52
    sub = (2^104 + 2^52) - hi                 exact (a multiple of 2^52 below 2^105)
    lo  = fma_rz(a, b, sub)                   = 2^52 + (ab mod 2^52), in [2^52, 2^53): exact
so the 52 mantissa bits of hi / lo ARE floor(ab / 2^52) / ab mod 2^52.  The kernel never converts: it adds the raw 64-bit
patterns of hi and lo into per-column uint64 accumulators (wrap-around arithmetic) and removes the statically known sum of
the exponent fields per column at the end.  Three FP64-pipe instructions and two 64-bit integer additions per limb product.
"""
from __future__ import annotations

P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
LIMBS, BITS = 8, 52
MASK = (1 << BITS) - 1
R = 1 << (LIMBS * BITS)
NPRIME = (-pow(P, -1, 1 << BITS)) & MASK          # -p^-1 mod 2^52
C_HI = 1 << 104
EXP_HI = (104 + 1023) << 52                       # exponent field of a double in [2^104, 2^105)
EXP_LO = (52 + 1023) << 52                        # exponent field of a double in [2^52, 2^53)
U64 = (1 << 64) - 1
P_LIMBS = [(P >> (BITS * i)) & MASK for i in range(LIMBS)]


def to_limbs(x: int):
    return [(x >> (BITS * i)) & MASK for i in range(LIMBS)]


def from_limbs(v) -> int:
    return sum(int(x) << (BITS * i) for i, x in enumerate(v))


def fma_rz(a: int, b: int, c: int) -> int:
    """a * b + c rounded toward zero to a 53-bit significand (the operands here are non-negative integers)."""
    s = a * b + c
    assert s >= 0
    drop = max(0, s.bit_length() - 53)
    return (s >> drop) << drop


def raw_bits(x: int) -> int:
    """IEEE-754 binary64 pattern of the positive integer x (x must be exactly representable)."""
    e = x.bit_length() - 1
    assert x >> max(0, e - 52) << max(0, e - 52) == x, "not representable"
    mant = (x << 52 >> e) & MASK if e <= 52 else (x >> (e - 52)) & MASK
    return ((e + 1023) << 52) | mant


def limb_product(a: int, b: int):
    """-> (raw pattern of hi, raw pattern of lo); 2 DFMA + 1 DADD on the FP64 pipe"""
    hi = fma_rz(a, b, C_HI)
    sub = (C_HI + (1 << 52)) - hi
    lo = fma_rz(a, b, sub)
    assert (1 << 52) <= lo < (1 << 53) and lo == a * b + sub          # exact
    return raw_bits(hi), raw_bits(lo)


def mont_mul(a_limbs, b_limbs, stats=None):
    """a b R^-1 mod p on 52-bit limbs; every limb product goes through limb_product, every accumulation is uint64 wrap-around"""
    col = [0] * (2 * LIMBS + 1)
    n_hi = [0] * (2 * LIMBS + 1)
    n_lo = [0] * (2 * LIMBS + 1)

    def acc(k, a, b):
        h, l = limb_product(a, b)
        col[k] = (col[k] + l) & U64; n_lo[k] += 1
        col[k + 1] = (col[k + 1] + h) & U64; n_hi[k + 1] += 1
        if stats is not None:
            stats["limb_products"] = stats.get("limb_products", 0) + 1

    for i in range(LIMBS):
        for j in range(LIMBS):
            acc(i + j, a_limbs[i], b_limbs[j])
    carry = 0
    for k in range(LIMBS):
        # column k is complete: strip the exponent fields, add the carry of the column below
        t = ((col[k] - n_hi[k] * EXP_HI - n_lo[k] * EXP_LO) & U64) + carry
        assert t < (1 << 64)
        q = ((t & MASK) * NPRIME) & MASK                       # low 52 bits: integer IMADs (or one more lo-product)
        col[k] = t; n_hi[k] = n_lo[k] = 0                       # from here on the column holds a plain integer
        for j in range(LIMBS):
            h, l = limb_product(q, P_LIMBS[j])
            if j == 0:
                col[k] = (col[k] + (l & MASK))                  # same column, already stripped: add the mantissa
            else:
                col[k + j] = (col[k + j] + l) & U64; n_lo[k + j] += 1
            col[k + j + 1] = (col[k + j + 1] + h) & U64; n_hi[k + j + 1] += 1
            if stats is not None:
                stats["limb_products"] = stats.get("limb_products", 0) + 1
        assert col[k] & MASK == 0
        carry = col[k] >> BITS
    out = []
    for k in range(LIMBS, 2 * LIMBS):
        t = ((col[k] - n_hi[k] * EXP_HI - n_lo[k] * EXP_LO) & U64) + carry
        assert t < (1 << 64)
        out.append(t & MASK)
        carry = t >> BITS
    top = ((col[2 * LIMBS] - n_hi[2 * LIMBS] * EXP_HI - n_lo[2 * LIMBS] * EXP_LO) & U64) + carry
    r = from_limbs(out) + (top << (BITS * LIMBS))
    if r >= P:
        r -= P
    assert r < P
    return to_limbs(r)


def mont_mul_r384(a_limbs, b_limbs):
    """a b 2^-384 mod p on the same 52-bit limbs: seven 52-bit reduction steps and one 20-bit step (7 * 52 + 20 = 384), so that the
    FP64-form field shares its Montgomery radix with the library's 12 x 32-bit form (fp.cuh) and values move between the two
    by re-packing bits only.  Same limb products, same accumulators; the result is shifted right by 20 bits at the end."""
    col = [0] * (2 * LIMBS + 1)
    bias = [0] * (2 * LIMBS + 1)

    def acc(k, a, b, skip_lo=False):
        h, l = limb_product(a, b)
        if not skip_lo:
            col[k] = (col[k] + l) & U64; bias[k] += EXP_LO
        col[k + 1] = (col[k + 1] + h) & U64; bias[k + 1] += EXP_HI
        return l & MASK

    for i in range(LIMBS):
        for j in range(LIMBS):
            acc(i + j, a_limbs[i], b_limbs[j])
    carry = 0
    for k in range(LIMBS):
        qbits = BITS if k < LIMBS - 1 else 20
        qmask = (1 << qbits) - 1
        t = ((col[k] - bias[k]) & U64) + carry
        assert t < (1 << 64)
        q = ((t & qmask) * NPRIME) & qmask
        for j in range(LIMBS):
            if j == 0:
                t += acc(k, q, P_LIMBS[0], skip_lo=True)
            else:
                acc(k + j, q, P_LIMBS[j])
        assert t & qmask == 0
        if k < LIMBS - 1:
            carry = t >> BITS
        else:
            col[k], bias[k], carry = t, 0, 0                     # the last column keeps its upper 32 bits
    wide, carry = 0, 0
    for k in range(LIMBS - 1, 2 * LIMBS + 1):
        t = ((col[k] - bias[k]) & U64) + carry
        assert t < (1 << 64)
        wide |= (t & MASK) << (BITS * (k - (LIMBS - 1)))
        carry = t >> BITS
    assert carry == 0
    r = wide >> 20
    if r >= P:
        r -= P
    assert r < P
    return to_limbs(r)


def budget():
    """pipe / issue accounting per product: returns a dict of cycles per SM sub-partition"""
    prods = 2 * LIMBS * LIMBS                                   # a*b and q*p
    fp64_ops = prods * 3 + 2 * LIMBS                            # + limb <-> double moves (one DADD each way)
    int_ops = prods * 4 + 8 * LIMBS                             # two 64-bit additions per product = 4 IADD3; strip/carry/q per column
    return {"limb_products": prods, "fp64_pipe_cycles": 2 * fp64_ops, "issue_slots": fp64_ops + int_ops,
            "imad_form_pipe_cycles": 1208, "imad_form_issue_slots": 302}


if __name__ == "__main__":
    import random
    rng = random.Random(1)
    for _ in range(200):
        a, b = rng.randrange(P), rng.randrange(P)
        got = from_limbs(mont_mul(to_limbs(a), to_limbs(b)))
        assert got == a * b * pow(R, -1, P) % P
    print("ok", budget())
