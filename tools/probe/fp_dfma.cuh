// fp_dfma.cuh -- PROBE (not product code, not measured yet): the BLS12-381 base-field Montgomery product carried by the FP64
// pipe, as specified and checked by tools/probe/dfma_model.py.  8 limbs of 52 bits held as doubles (exact integers < 2^52),
// Montgomery radix R = 2^416.  Per limb product: hi = fma.rz(a, b, 2^104), lo = fma.rz(a, b, (2^104 + 2^52) - hi); the 52
// mantissa bits of hi / lo are floor(ab / 2^52) / ab mod 2^52.  The raw 64-bit patterns are accumulated in wrap-around uint64
// columns; the statically known sum of the exponent fields of a column (`bias`, constant-folded after unrolling) is removed
// when the column is consumed.  Compiles for sm_100a (instruction mix: tools/probe/README_dfma.txt) and for the host
// (tests/test_dfma_model.py, fma() under FE_TOWARDZERO).
#pragma once
#include <stdint.h>
#include <cfenv>
#include <cmath>
#include <cstring>
#if defined(__CUDACC__)
#define DF_D __host__ __device__ __forceinline__
#else
#define DF_D inline
#endif

namespace dfma {

constexpr int L = 8;
constexpr uint64_t MASK = (1ull << 52) - 1;
constexpr uint64_t EXP_HI = (uint64_t)(104 + 1023) << 52;
constexpr uint64_t EXP_LO = (uint64_t)(52 + 1023) << 52;
constexpr uint64_t NPRIME = 0x3fffcfffcfffdull;                 // -p^-1 mod 2^52
constexpr double C_HI = 20282409603651670423947251286016.0;      // 2^104
constexpr double C_HI_PLUS = 20282409603651674927546878656512.0; // 2^104 + 2^52
constexpr double TWO52 = 4503599627370496.0;

DF_D constexpr uint64_t p_limb(int i) {
  return i == 0 ? 0xeffffffffaaabull : i == 1 ? 0xfeb153ffffb9full : i == 2 ? 0x6b0f6241eabffull : i == 3 ? 0x12bf6730d2a0full
       : i == 4 ? 0x764774b84f385ull : i == 5 ? 0x1ba7b6434bacdull : i == 6 ? 0x1ea397fe69a4bull : 0x1a011ull;
}

// host code (tests, the probe's self-check) runs under FE_TOWARDZERO, set by the caller
DF_D double fma_rz(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rz(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}
DF_D uint64_t bits(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t r; std::memcpy(&r, &x, 8); return r;
#endif
}
DF_D double from_bits(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)x);
#else
  double r; std::memcpy(&r, &x, 8); return r;
#endif
}

// exact double of an integer below 2^52: one integer OR + one FP64 add
DF_D double limb_to_double(uint64_t x) { return from_bits(x | EXP_LO) - TWO52; }

struct Fp { double v[L]; };

// 2 DFMA + 1 DADD
DF_D void limb_product(double a, double b, uint64_t& hi, uint64_t& lo) {
  const double h = fma_rz(a, b, C_HI);
  const double sub = C_HI_PLUS - h;
  const double l = fma_rz(a, b, sub);
  hi = bits(h);
  lo = bits(l);
}

// r = a b R^-1 mod p
DF_D void mont_mul(Fp& r, const Fp& a, const Fp& b) {
  uint64_t col[2 * L + 1], bias[2 * L + 1];
#pragma unroll
  for (int k = 0; k < 2 * L + 1; k++) { col[k] = 0; bias[k] = 0; }
#pragma unroll
  for (int i = 0; i < L; i++) {
#pragma unroll
    for (int j = 0; j < L; j++) {
      uint64_t h, l;
      limb_product(a.v[i], b.v[j], h, l);
      col[i + j] += l; bias[i + j] += EXP_LO;
      col[i + j + 1] += h; bias[i + j + 1] += EXP_HI;
    }
  }
  uint64_t carry = 0;
#pragma unroll
  for (int k = 0; k < L; k++) {
    uint64_t t = col[k] - bias[k] + carry;                       // the column as a plain integer (< 2^57)
    const uint64_t q = ((t & MASK) * NPRIME) & MASK;             // low 52 bits of a 52 x 52 product: integer multiply
    const double qd = limb_to_double(q);
#pragma unroll
    for (int j = 0; j < L; j++) {
      uint64_t h, l;
      limb_product(qd, limb_to_double(p_limb(j)), h, l);         // limb_to_double of a constant folds
      if (j == 0) t += l & MASK;
      else { col[k + j] += l; bias[k + j] += EXP_LO; }
      col[k + j + 1] += h; bias[k + j + 1] += EXP_HI;
    }
    carry = t >> 52;                                             // t mod 2^52 == 0 by the choice of q
  }
  uint64_t out[L];
#pragma unroll
  for (int k = 0; k < L; k++) {
    const uint64_t t = col[L + k] - bias[L + k] + carry;
    out[k] = t & MASK;
    carry = t >> 52;
  }
  // a, b < p  =>  result < 2p: col[2L] and the last carry are zero (R = 2^416 > 4p); one conditional subtraction of p
  uint64_t d[L];
  uint64_t borrow = 0;
#pragma unroll
  for (int k = 0; k < L; k++) {
    const uint64_t s = out[k] - p_limb(k) - borrow;
    d[k] = s & MASK;
    borrow = (s >> 63) & 1;
  }
#pragma unroll
  for (int k = 0; k < L; k++) r.v[k] = limb_to_double(borrow ? out[k] : d[k]);
}

// r = a b 2^-384 mod p: the library's Montgomery radix (fp.cuh, 12 x 32-bit limbs), so that a value moves between the two forms
// by re-packing bits only (from_u32 / to_u32 below).  Seven 52-bit reduction steps and one 20-bit step (7 * 52 + 20 = 384); the
// columns from the seventh upward are shifted right by 20 bits at the end.  Same 128 limb products as mont_mul.
DF_D void mont_mul384(Fp& r, const Fp& a, const Fp& b) {
  uint64_t col[2 * L + 1], bias[2 * L + 1];
#pragma unroll
  for (int k = 0; k < 2 * L + 1; k++) { col[k] = 0; bias[k] = 0; }
#pragma unroll
  for (int i = 0; i < L; i++) {
#pragma unroll
    for (int j = 0; j < L; j++) {
      uint64_t h, l;
      limb_product(a.v[i], b.v[j], h, l);
      col[i + j] += l; bias[i + j] += EXP_LO;
      col[i + j + 1] += h; bias[i + j + 1] += EXP_HI;
    }
  }
  uint64_t carry = 0;
#pragma unroll
  for (int k = 0; k < L; k++) {
    const uint64_t qmask = k < L - 1 ? MASK : ((1ull << 20) - 1);
    uint64_t t = col[k] - bias[k] + carry;
    const uint64_t q = ((t & qmask) * NPRIME) & qmask;
    const double qd = limb_to_double(q);
#pragma unroll
    for (int j = 0; j < L; j++) {
      uint64_t h, l;
      limb_product(qd, limb_to_double(p_limb(j)), h, l);
      if (j == 0) t += l & MASK;
      else { col[k + j] += l; bias[k + j] += EXP_LO; }
      col[k + j + 1] += h; bias[k + j + 1] += EXP_HI;
    }
    if (k < L - 1) carry = t >> 52;
    else { col[k] = t; bias[k] = 0; carry = 0; }                 // low 20 bits are zero, the rest belongs to the result
  }
  uint64_t w[L + 2];                                             // columns 7 .. 16 as 52-bit limbs
#pragma unroll
  for (int k = 0; k < L + 2; k++) {
    const uint64_t t = col[L - 1 + k] - bias[L - 1 + k] + carry;
    w[k] = t & MASK;
    carry = t >> 52;
  }
  uint64_t out[L];
#pragma unroll
  for (int k = 0; k < L; k++) out[k] = ((w[k] >> 20) | (w[k + 1] << 32)) & MASK;
  uint64_t d[L];
  uint64_t borrow = 0;
#pragma unroll
  for (int k = 0; k < L; k++) {
    const uint64_t s = out[k] - p_limb(k) - borrow;
    d[k] = s & MASK;
    borrow = (s >> 63) & 1;
  }
#pragma unroll
  for (int k = 0; k < L; k++) r.v[k] = limb_to_double(borrow ? out[k] : d[k]);
}

// 12 x 32-bit little-endian limbs (the library's Fp<Bls381Fp>::v) <-> 8 x 52-bit double limbs: the same integer, re-packed
DF_D void from_u32(Fp& r, const uint32_t* v) {
#pragma unroll
  for (int i = 0; i < L; i++) {
    const int bit = 52 * i, w = bit >> 5, sh = bit & 31;
    uint64_t x = (uint64_t)v[w] >> sh;
    if (w + 1 < 12) x |= (uint64_t)v[w + 1] << (32 - sh);
    if (w + 2 < 12 && sh > 12) x |= (uint64_t)v[w + 2] << (64 - sh);
    r.v[i] = limb_to_double(x & MASK);
  }
}
DF_D void to_u32(uint32_t* v, const Fp& a) {
  uint64_t x[L];
#pragma unroll
  for (int i = 0; i < L; i++) x[i] = bits(a.v[i] + TWO52) & MASK;
#pragma unroll
  for (int w = 0; w < 12; w++) {
    const int bit = 32 * w, i = bit / 52, sh = bit % 52;
    uint64_t y = x[i] >> sh;
    if (sh > 20 && i + 1 < L) y |= x[i + 1] << (52 - sh);
    v[w] = (uint32_t)y;
  }
}

}  // namespace dfma
