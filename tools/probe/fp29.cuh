// fp29.cuh -- EXPERIMENT, not part of the library: BLS12-381 base field in an unsaturated radix (14 limbs of 29 bits,
// Montgomery R = 2^406) with carry-free 64-bit column accumulation.
//
// Hypothesis tested: "plain IMAD.WIDE.U32 issues twice as fast as the carry-chained IMAD.WIDE.U32.X of the saturated
// 12 x 32-bit product".  RESULT (B200, tools/probe/imad_probe.cu, fpmul_probe.cu): false -- every IMAD.WIDE holds the
// FMA-heavy pipe 4 cycles per warp instruction, with or without carry, register or immediate operands; this product
// (392 IMAD.WIDE + ~360 ALU instructions) sustains 2.0e10 products/s against 3.0e10 for fp.cuh (288 IMAD.WIDE-class).
// The code is correct (checked on the host against Python integers) and stays here as the measured dead end.
//
// Representation invariants ("normalised"): limbs 0..12 < 2^29, limb 13 holds everything above bit 377; the VALUE may
// exceed p:  products return values < 2p, sums/differences add up; any input of a product must stay below 2^9 p (then
// a b / R < p and the result is again < 2p).  f_sub(a, b) = a - b + 32p needs b < 32p.  Exact zero/equality tests
// (only needed for the exceptional cases of the group law) reduce first.  Invariants are asserted in the host
// emulation build (B2K_HOST_EMUL), which runs every kernel body on the CPU in the tests.
#pragma once
#include "fp.cuh"
#include "constants.cuh"
#include "fp29_constants.cuh"
#if defined(B2K_HOST_EMUL) && !defined(__CUDACC__)
#include <cassert>
#define B2K_ASSERT29(x) assert(x)
#else
#define B2K_ASSERT29(x)
#endif

namespace b2k {

template <class C>
struct Fp29 { uint32_t v[14]; };

constexpr uint32_t M29 = (1u << 29) - 1u;

template <class C>
B2K_D void fp29_check(const Fp29<C>& a, uint32_t top_bound) {
#if defined(B2K_HOST_EMUL) && !defined(__CUDACC__)
  for (int j = 0; j < 13; j++) B2K_ASSERT29(a.v[j] <= M29);
  B2K_ASSERT29(a.v[13] <= top_bound);
#else
  (void)a; (void)top_bound;
#endif
}
// top limb of k p is 13 k (+ carry): bounds used by the assertions
constexpr uint32_t TOP_2P = 27, TOP_32P = 415, TOP_MULIN = 13u * 512u + 8u;

// carry-propagate limbs that may have grown up to 32 bits
template <class C>
B2K_D void fp29_normalise(Fp29<C>& r, const uint32_t* d) {
  uint32_t carry = 0;
#pragma unroll
  for (int j = 0; j < 13; j++) {
    const uint32_t t = d[j] + carry;          // d[j] < 2^31 + small: no wrap
    r.v[j] = t & M29;
    carry = t >> 29;
  }
  r.v[13] = d[13] + carry;
}

template <class C> B2K_D void fp29_set_zero(Fp29<C>& r) {
#pragma unroll
  for (int j = 0; j < 14; j++) r.v[j] = 0;
}
template <class C> B2K_D void fp29_set_one(Fp29<C>& r) {
#pragma unroll
  for (int j = 0; j < 14; j++) r.v[j] = C::r1(j);
}

template <class C>
B2K_D void fp29_add(Fp29<C>& r, const Fp29<C>& a, const Fp29<C>& b) {
  uint32_t d[14];
#pragma unroll
  for (int j = 0; j < 14; j++) d[j] = a.v[j] + b.v[j];
  fp29_normalise(r, d);
  fp29_check(r, TOP_MULIN);
}

// a - b + 32p   (b < 32p)
template <class C>
B2K_D void fp29_sub(Fp29<C>& r, const Fp29<C>& a, const Fp29<C>& b) {
  fp29_check(b, TOP_32P);
  uint32_t d[14];
#pragma unroll
  for (int j = 0; j < 14; j++) d[j] = a.v[j] + (C::sub_c(j) - b.v[j]);
  fp29_normalise(r, d);
  fp29_check(r, TOP_MULIN);
}
template <class C>
B2K_D void fp29_neg(Fp29<C>& r, const Fp29<C>& a) {
  fp29_check(a, TOP_32P);
  uint32_t d[14];
#pragma unroll
  for (int j = 0; j < 14; j++) d[j] = C::sub_c(j) - a.v[j];
  fp29_normalise(r, d);
}

// Montgomery reduction of the 28 column sums t[] (each < 2^63) and carry-normalised output < 2p
template <class C>
B2K_D void fp29_reduce(Fp29<C>& r, uint64_t* t) {
#pragma unroll
  for (int i = 0; i < 14; i++) {
    const uint32_t m = ((uint32_t)t[i] * C::PINV) & M29;
#pragma unroll
    for (int j = 0; j < 14; j++) t[i + j] += (uint64_t)m * C::mod(j);
    t[i + 1] += t[i] >> 29;                   // low 29 bits are zero by construction
  }
#pragma unroll
  for (int k = 0; k < 13; k++) {
    r.v[k] = (uint32_t)t[14 + k] & M29;
    t[15 + k] += t[14 + k] >> 29;
  }
  r.v[13] = (uint32_t)t[27];
  fp29_check(r, TOP_2P);
}

template <class C>
B2K_D void fp29_mul(Fp29<C>& r, const Fp29<C>& a, const Fp29<C>& b) {
  fp29_check(a, TOP_MULIN); fp29_check(b, TOP_MULIN);
  uint64_t t[28];
#pragma unroll
  for (int k = 0; k < 28; k++) t[k] = 0;
#pragma unroll
  for (int i = 0; i < 14; i++)
#pragma unroll
    for (int j = 0; j < 14; j++) t[i + j] += (uint64_t)a.v[i] * b.v[j];
  fp29_reduce(r, t);
}

template <class C>
B2K_D void fp29_sqr(Fp29<C>& r, const Fp29<C>& a) {
  fp29_check(a, TOP_MULIN);
  uint64_t t[28];
#pragma unroll
  for (int k = 0; k < 28; k++) t[k] = 0;
#pragma unroll
  for (int i = 0; i < 14; i++) {
    t[2 * i] += (uint64_t)a.v[i] * a.v[i];
    const uint32_t a2 = a.v[i] << 1;          // < 2^30 for limbs < 2^29; the top limb (< 2^13) stays small too
#pragma unroll
    for (int j = i + 1; j < 14; j++) t[i + j] += (uint64_t)a2 * a.v[j];
  }
  fp29_reduce(r, t);
}

// out-of-line copies for cold paths
template <class C> B2K_NI void fp29_mul_c(Fp29<C>& r, const Fp29<C>& a, const Fp29<C>& b) { fp29_mul(r, a, b); }

// value < 2p (a product's output): is it 0 mod p ?
template <class C>
B2K_D bool fp29_is_zero_lt2p(const Fp29<C>& a) {
  uint32_t z = 0, e = 0;
#pragma unroll
  for (int j = 0; j < 14; j++) { z |= a.v[j]; e |= a.v[j] ^ C::mod(j); }
  return z == 0 || e == 0;
}
// exact test for any normalised value below 2^9 p: one reduction (a * 1 / R  is  0 mod p  iff  a is)
template <class C>
B2K_D bool fp29_is_zero(const Fp29<C>& a) {
  uint64_t t[28];
#pragma unroll
  for (int k = 0; k < 28; k++) t[k] = k < 14 ? (uint64_t)a.v[k] : 0;
  Fp29<C> r;
  fp29_reduce(r, t);
  return fp29_is_zero_lt2p(r);
}
template <class C>
B2K_D bool fp29_eq(const Fp29<C>& a, const Fp29<C>& b) {
  Fp29<C> d;
  fp29_sub(d, a, b);
  return fp29_is_zero(d);
}

// ---- conversions with the saturated 12 x 32-bit Montgomery form (R = 2^384) of fp.cuh -----------------------------
// re-split 12 x 32 bits into 14 x 29 bits (pure bit movement; the integer is unchanged)
B2K_D void limbs32_to_29(uint32_t* o, const uint32_t* v) {
#pragma unroll
  for (int j = 0; j < 14; j++) {
    const int bit = 29 * j, w = bit >> 5, s = bit & 31;
    uint64_t two = (uint64_t)v[w];
    if (w + 1 < 12) two |= (uint64_t)v[w + 1] << 32;
    o[j] = (uint32_t)(two >> s) & M29;
  }
}
B2K_D void limbs29_to_32(uint32_t* o, const uint32_t* v) {   // v normalised and < 2^384
#pragma unroll
  for (int w = 0; w < 12; w++) {
    const int bit = 32 * w, j = bit / 29, s = bit - 29 * j;   // limb j bit s is bit 0 of word w
    uint64_t acc = (uint64_t)v[j] >> s;
    acc |= (uint64_t)v[j + 1] << (29 - s);
    if (j + 2 < 14) acc |= (uint64_t)v[j + 2] << (58 - s);
    o[w] = (uint32_t)acc;
  }
}

template <class C29, class C32>
B2K_D void fp29_from_fp32(Fp29<C29>& r, const Fp<C32>& a_mont384) {
  Fp29<C29> s, f;
  limbs32_to_29(s.v, a_mont384.v);
#pragma unroll
  for (int j = 0; j < 14; j++) f.v[j] = C29::from_r384(j);
  fp29_mul(r, s, f);
}
// canonical (< p) Montgomery-384 value from any normalised Fp29 below 2^9 p
template <class C29, class C32>
B2K_D void fp29_to_fp32(Fp<C32>& r, const Fp29<C29>& a) {
  Fp29<C29> g, t;
#pragma unroll
  for (int j = 0; j < 14; j++) g.v[j] = C29::to_r384(j);
  fp29_mul(t, a, g);                           // < 2p
  // subtract p once if t >= p
  uint32_t d[14];
  int32_t br = 0;
#pragma unroll
  for (int j = 0; j < 13; j++) {
    const int32_t x = (int32_t)t.v[j] - (int32_t)C29::mod(j) + br;
    d[j] = (uint32_t)x & M29;
    br = x >> 29;                              // 0 or -1
  }
  const int32_t top = (int32_t)t.v[13] - (int32_t)C29::mod(13) + br;
  d[13] = (uint32_t)top;
  const bool ge = top >= 0;
  uint32_t c[14];
#pragma unroll
  for (int j = 0; j < 14; j++) c[j] = ge ? d[j] : t.v[j];
  limbs29_to_32(r.v, c);
}

// ---- the generic f_* interface the group law templates use ----------------------------------------------------------
template <class C> B2K_D void f_add(Fp29<C>& r, const Fp29<C>& a, const Fp29<C>& b) { fp29_add(r, a, b); }
template <class C> B2K_D void f_sub(Fp29<C>& r, const Fp29<C>& a, const Fp29<C>& b) { fp29_sub(r, a, b); }
template <class C> B2K_D void f_mul(Fp29<C>& r, const Fp29<C>& a, const Fp29<C>& b) { fp29_mul(r, a, b); }
template <class C> B2K_D void f_sqr(Fp29<C>& r, const Fp29<C>& a) { fp29_sqr(r, a); }
template <class C> B2K_D void f_neg(Fp29<C>& r, const Fp29<C>& a) { fp29_neg(r, a); }
template <class C> B2K_D void f_dbl(Fp29<C>& r, const Fp29<C>& a) { fp29_add(r, a, a); }
template <class C> B2K_D bool f_is_zero(const Fp29<C>& a) { return fp29_is_zero(a); }
template <class C> B2K_D bool f_eq(const Fp29<C>& a, const Fp29<C>& b) { return fp29_eq(a, b); }
template <class C> B2K_D void f_set_zero(Fp29<C>& r) { fp29_set_zero(r); }
template <class C> B2K_D void f_set_one(Fp29<C>& r) { fp29_set_one(r); }

}  // namespace b2k
