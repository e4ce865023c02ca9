// ec_dfma.cuh -- PROBE (next round's dual-pipe bucket arithmetic, not product code): field add/sub and the mixed XYZZ addition of
// the MSM accumulate pass (ec.cuh: xyzz_add_mixed, 8 M + 2 S) on the FP64-pipe field of fp_dfma.cuh.  Limbs are doubles holding
// exact integers < 2^52, values canonical (< p) in Montgomery form with R = 2^416.  Additions and subtractions go through the
// integer view of the limbs (one FP64 add + one mask per limb each way): 7 of them against 10 products per point addition.
#pragma once
#include "fp_dfma.cuh"

namespace dfma {

// integer view: bits(x + 2^52) has the integer in its mantissa (x < 2^52 exact)
DF_D uint64_t limb_to_int(double x) { return bits(x + TWO52) & MASK; }

DF_D bool fp_is_zero(const Fp& a) {
  double o = 0.0;
#pragma unroll
  for (int i = 0; i < L; i++) o += a.v[i];          // limbs are non-negative: the sum is zero iff all are
  return o == 0.0;
}

DF_D void fp_add(Fp& r, const Fp& a, const Fp& b) {
  uint64_t s[L], d[L], carry = 0, borrow = 0;
#pragma unroll
  for (int i = 0; i < L; i++) {
    const uint64_t t = limb_to_int(a.v[i]) + limb_to_int(b.v[i]) + carry;
    s[i] = t & MASK; carry = t >> 52;
  }
#pragma unroll
  for (int i = 0; i < L; i++) {                      // a + b < 2p < 2^416: no carry out of the top limb
    const uint64_t t = s[i] - p_limb(i) - borrow;
    d[i] = t & MASK; borrow = (t >> 63) & 1;
  }
#pragma unroll
  for (int i = 0; i < L; i++) r.v[i] = limb_to_double(borrow ? s[i] : d[i]);
}

DF_D void fp_sub(Fp& r, const Fp& a, const Fp& b) {
  uint64_t d[L], e[L], borrow = 0, carry = 0;
#pragma unroll
  for (int i = 0; i < L; i++) {
    const uint64_t t = limb_to_int(a.v[i]) - limb_to_int(b.v[i]) - borrow;
    d[i] = t & MASK; borrow = (t >> 63) & 1;
  }
#pragma unroll
  for (int i = 0; i < L; i++) {                      // + p when the difference went negative
    const uint64_t t = d[i] + p_limb(i) + carry;
    e[i] = t & MASK; carry = t >> 52;
  }
#pragma unroll
  for (int i = 0; i < L; i++) r.v[i] = limb_to_double(borrow ? e[i] : d[i]);
}

struct Affine { Fp x, y; };            // infinity is not representable here: the caller skips such operands
struct Xyzz { Fp X, Y, ZZ, ZZZ; };     // x = X / ZZ, y = Y / ZZZ; infinity: ZZ = 0

// acc += p.  Returns 0 = done, 1 = acc was infinity (acc := p done here), 2 = same point (the caller doubles on the rare path),
// 3 = opposite points (acc := infinity done here).  Same case analysis as ec.cuh: xyzz_add_mixed.
// R384 = false: products by mont_mul (radix 2^416); true: mont_mul384 (the library's radix 2^384: operands are re-packed library values)
template <bool R384 = false>
DF_D void mul_sel(Fp& r, const Fp& a, const Fp& b) {
  if (R384) mont_mul384(r, a, b);
  else mont_mul(r, a, b);
}
template <bool R384 = false>
DF_D int xyzz_madd(Xyzz& acc, const Affine& p, const Fp& one_mont) {
  if (fp_is_zero(acc.ZZ)) { acc.X = p.x; acc.Y = p.y; acc.ZZ = one_mont; acc.ZZZ = one_mont; return 1; }
  Fp U2, S2, P, Rr, PP, PPP, Q, t, X3, Y3;
  mul_sel<R384>(U2, p.x, acc.ZZ);
  mul_sel<R384>(S2, p.y, acc.ZZZ);
  fp_sub(P, U2, acc.X);
  fp_sub(Rr, S2, acc.Y);
  if (fp_is_zero(P)) {
    if (fp_is_zero(Rr)) return 2;
#pragma unroll
    for (int i = 0; i < L; i++) acc.ZZ.v[i] = 0.0;
    acc.ZZZ = acc.ZZ;
    return 3;
  }
  mul_sel<R384>(PP, P, P);
  mul_sel<R384>(PPP, P, PP);
  mul_sel<R384>(Q, acc.X, PP);
  mul_sel<R384>(X3, Rr, Rr);
  fp_sub(X3, X3, PPP);
  fp_sub(X3, X3, Q);
  fp_sub(X3, X3, Q);
  fp_sub(t, Q, X3);
  mul_sel<R384>(Y3, Rr, t);
  mul_sel<R384>(t, acc.Y, PPP);
  fp_sub(Y3, Y3, t);
  mul_sel<R384>(t, acc.ZZ, PP);
  acc.ZZ = t;
  mul_sel<R384>(t, acc.ZZZ, PPP);
  acc.ZZZ = t;
  acc.X = X3;
  acc.Y = Y3;
  return 0;
}

}  // namespace dfma
