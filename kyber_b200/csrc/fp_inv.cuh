// fp_inv.cuh -- variable-time modular inversion by the binary extended Euclid ("almost Montgomery inverse",
// Kaliski 1995): ~1.4 bitlen(p) iterations of shifts and subtractions instead of a ~bitlen(p)-squaring Fermat
// ladder.  Used where ONE thread inverts (the serial tail of the MSM); data-dependent branches make it
// unsuitable for warp-wide batches (those use Fermat or a block-level batch inversion).
//
// Input a (Montgomery form aR), output a^-1 R:
//   phase 1: x = (aR)^-1 2^k mod p, n <= k <= 2n          (u, v, r, s with r, s < 2p)
//   phase 2: k halvings mod p  ->  a^-1 R^-1
//   phase 3: one Montgomery product with R^3 mod p        ->  a^-1 R
// The inverse of 0 is 0 (as with the Fermat version).
#pragma once
#include "fp.cuh"

namespace b2k {
namespace detail {

template <int N> B2K_D bool mp_is_zero(const uint32_t* a) { uint32_t o = 0; for (int j = 0; j < N; j++) o |= a[j]; return o == 0; }
template <int N> B2K_D bool mp_gt(const uint32_t* a, const uint32_t* b) {      // a > b
  for (int j = N - 1; j >= 0; j--) {
    if (a[j] > b[j]) return true;
    if (a[j] < b[j]) return false;
  }
  return false;
}
template <int N> B2K_D void mp_shr1(uint32_t* a) {
  for (int j = 0; j < N - 1; j++) a[j] = (a[j] >> 1) | (a[j + 1] << 31);
  a[N - 1] >>= 1;
}
template <int N> B2K_D void mp_shl1(uint32_t* a) {
  for (int j = N - 1; j > 0; j--) a[j] = (a[j] << 1) | (a[j - 1] >> 31);
  a[0] <<= 1;
}
template <int N> B2K_D void mp_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {   // r = a - b (a >= b)
  r[0] = ptx::sub_cc(a[0], b[0]);
  for (int j = 1; j < N - 1; j++) r[j] = ptx::subc_cc(a[j], b[j]);
  r[N - 1] = ptx::subc(a[N - 1], b[N - 1]);
}
template <int N> B2K_D void mp_add(uint32_t* r, const uint32_t* a, const uint32_t* b) {   // r = a + b (no overflow)
  r[0] = ptx::add_cc(a[0], b[0]);
  for (int j = 1; j < N - 1; j++) r[j] = ptx::addc_cc(a[j], b[j]);
  r[N - 1] = ptx::addc(a[N - 1], b[N - 1]);
}

}  // namespace detail

template <class C>
B2K_NI void fp_inv_vartime(Fp<C>& out, const Fp<C>& a) {
  constexpr int N = C::N;
  using namespace detail;
  uint32_t u[N], v[N], r[N], s[N], p[N];
  for (int j = 0; j < N; j++) { p[j] = C::mod(j); u[j] = p[j]; v[j] = a.v[j]; r[j] = 0; s[j] = 0; }
  s[0] = 1;
  if (mp_is_zero<N>(v)) { fp_set_zero(out); return; }
  int k = 0;
  while (!mp_is_zero<N>(v)) {
    if (!(u[0] & 1)) { mp_shr1<N>(u); mp_shl1<N>(s); }
    else if (!(v[0] & 1)) { mp_shr1<N>(v); mp_shl1<N>(r); }
    else if (mp_gt<N>(u, v)) { mp_sub<N>(u, u, v); mp_shr1<N>(u); mp_add<N>(r, r, s); mp_shl1<N>(s); }
    else { mp_sub<N>(v, v, u); mp_shr1<N>(v); mp_add<N>(s, s, r); mp_shl1<N>(r); }
    k++;
  }
  if (!mp_gt<N>(p, r)) mp_sub<N>(r, r, p);     // r >= p
  mp_sub<N>(r, p, r);                           // r = p - r = (aR)^-1 2^k
  for (int i = 0; i < k; i++) {                 // divide by 2^k mod p
    if (r[0] & 1) mp_add<N>(r, r, p);           // r < p and p < 2^(32N-1): no overflow
    mp_shr1<N>(r);
  }
  Fp<C> x, r3;
  for (int j = 0; j < N; j++) { x.v[j] = r[j]; r3.v[j] = C::r3(j); }
  fp_mul_c(out, x, r3);
}

// ------------------------------------------------------------------------------------------------------------------
// Branch-free inversion for warp-wide use: binary GCD on 64-bit approximations (Pornin, "Optimized Binary GCD for
// Modular Inversion", 2020, algorithm 2, restated for 32-bit limbs).  Every lane runs the same instruction stream
// (masks, no data-dependent branch except the early exit when a lane has converged), so 32 inversions cost one
// inversion's issue slots -- unlike fp_inv_vartime, whose four-way branch serialises inside a warp.
//   state: a, b >= 0 (N limbs), u, v in [0, p)  with  a = u y, b = v y (mod p);  start (y, p, 1, 0)
//   outer round: 30 divsteps on (low 31 bits | top 33 bits) of a and b give a 2x2 matrix (f0 g0; f1 g1), |entries| <= 2^30;
//                (a, b) <- |(f0 a + g0 b, f1 a + g1 b)| / 2^30   (exact);   (u, v) <- the same combination / 2^30 mod p
//   after ceil((2 bitlen(p) - 1) / 30) rounds a = 0, b = 1 and v = y^-1.  Extra rounds are the identity.
// Input and output in Montgomery form (aR -> a^-1 R: one product with R^3); the inverse of 0 is 0.
// Used by the affine pair-tree rounds of the MSM (msm_affine.cuh), where each thread inverts one running product.
namespace detail {

// w[0..N+1] = sa (a fa) + sb (b fb) in two's complement, sa/sb = -1 when ma/mb is all-ones
template <int N>
B2K_D void mp_lin2_signed(uint32_t* w, const uint32_t* a, uint32_t fa, uint32_t ma, const uint32_t* b, uint32_t fb, uint32_t mb) {
  uint64_t ca = 0, cb = 0, na = ma & 1u, nb = mb & 1u, cs = 0;
#pragma unroll
  for (int j = 0; j < N + 2; j++) {
    uint32_t pa = 0, pb = 0;
    if (j < N) {
      uint64_t t = (uint64_t)a[j] * fa + ca; pa = (uint32_t)t; ca = t >> 32;
      uint64_t s = (uint64_t)b[j] * fb + cb; pb = (uint32_t)s; cb = s >> 32;
    } else if (j == N) { pa = (uint32_t)ca; pb = (uint32_t)cb; }
    uint64_t xa = (uint64_t)(pa ^ ma) + na; na = xa >> 32;
    uint64_t xb = (uint64_t)(pb ^ mb) + nb; nb = xb >> 32;
    uint64_t s2 = (uint64_t)(uint32_t)xa + (uint32_t)xb + cs; w[j] = (uint32_t)s2; cs = s2 >> 32;
  }
}

// r = (x fx + y fy + t p) / 2^30 mod p  for x, y <= p and fx + fy <= 2^30;  t makes the sum divisible by 2^30
template <class C>
B2K_D void mp_lin2_mont30(uint32_t* r, const uint32_t* x, uint32_t fx, const uint32_t* y, uint32_t fy) {
  constexpr int N = C::N;
  uint32_t w[N + 1];
  uint64_t cx = 0, cy = 0, cs = 0;
#pragma unroll
  for (int j = 0; j < N; j++) {
    uint64_t t = (uint64_t)x[j] * fx + cx; cx = t >> 32;
    uint64_t s = (uint64_t)y[j] * fy + cy; cy = s >> 32;
    uint64_t z = (uint64_t)(uint32_t)t + (uint32_t)s + cs; w[j] = (uint32_t)z; cs = z >> 32;
  }
  w[N] = (uint32_t)(cx + cy + cs);
  const uint32_t t30 = (w[0] * C::M0) & 0x3fffffffu;         // M0 = -p^-1 mod 2^32
  uint64_t c = 0;
#pragma unroll
  for (int j = 0; j < N; j++) {
    uint64_t z = (uint64_t)C::mod(j) * t30 + w[j] + c; w[j] = (uint32_t)z; c = z >> 32;
  }
  w[N] += (uint32_t)c;
#pragma unroll
  for (int j = 0; j < N; j++) r[j] = (w[j] >> 30) | (w[j + 1] << 2);
  fp_reduce_once<C>(r);
}

}  // namespace detail

template <class C>
B2K_NI void fp_inv_bingcd(Fp<C>& out, const Fp<C>& x) {
  constexpr int N = C::N;
  static_assert(C::BITS <= 32 * C::N - 1, "the modulus must leave the top bit of the limb vector free");
  constexpr int ROUNDS = (2 * C::BITS - 1 + 29) / 30;
  using namespace detail;
  uint32_t a[N], b[N], u[N], v[N];
#pragma unroll
  for (int j = 0; j < N; j++) { a[j] = x.v[j]; b[j] = C::mod(j); u[j] = 0; v[j] = 0; }
  u[0] = 1;
  for (int it = 0; it < ROUNDS; it++) {
    if (mp_is_zero<N>(a)) break;                   // converged (or input 0): further rounds change nothing
    // ---- 64-bit approximations: exact when both fit 64 bits, else low 31 bits | top 33 bits of the longer one
    uint32_t ha = a[1], ma_ = a[0], la = 0, hb = b[1], mb_ = b[0], lb = 0;
    bool wide = false;
#pragma unroll
    for (int j = 2; j < N; j++) {
      if ((a[j] | b[j]) != 0) { ha = a[j]; ma_ = a[j - 1]; la = a[j - 2]; hb = b[j]; mb_ = b[j - 1]; lb = b[j - 2]; wide = true; }
    }
    uint64_t A = ((uint64_t)ha << 32) | ma_, Bv = ((uint64_t)hb << 32) | mb_;
    if (wide) {
      uint32_t top = ha | hb;                      // non-zero
      int sh = 0;
      while (!((top << sh) & 0x80000000u)) sh++;   // count leading zeros (<= 31), same code on host and device
      if (sh) { A = (A << sh) | (la >> (32 - sh)); Bv = (Bv << sh) | (lb >> (32 - sh)); }
      A = ((A >> 31) << 31) | (a[0] & 0x7fffffffu);
      Bv = ((Bv >> 31) << 31) | (b[0] & 0x7fffffffu);
    }
    uint32_t f0 = 1, g0 = 0, f1 = 0, g1 = 1;       // two's complement int32
#pragma unroll 5
    for (int i = 0; i < 30; i++) {
      const uint32_t odd = 0u - (uint32_t)(A & 1u);
      const uint32_t sw = odd & (0u - (uint32_t)(A < Bv));
      const uint64_t sw64 = ((uint64_t)sw << 32) | sw, odd64 = ((uint64_t)odd << 32) | odd;
      const uint64_t t = (A ^ Bv) & sw64; A ^= t; Bv ^= t;
      const uint32_t tf = (f0 ^ f1) & sw; f0 ^= tf; f1 ^= tf;
      const uint32_t tg = (g0 ^ g1) & sw; g0 ^= tg; g1 ^= tg;
      A -= Bv & odd64; f0 -= f1 & odd; g0 -= g1 & odd;
      A >>= 1; f1 <<= 1; g1 <<= 1;
    }
    // ---- (a, b) <- |f a + g b| / 2^30, signs folded into the matrix rows
    uint32_t mf0 = 0u - (f0 >> 31), mg0 = 0u - (g0 >> 31), mf1 = 0u - (f1 >> 31), mg1 = 0u - (g1 >> 31);
    uint32_t af0 = (f0 ^ mf0) - mf0, ag0 = (g0 ^ mg0) - mg0, af1 = (f1 ^ mf1) - mf1, ag1 = (g1 ^ mg1) - mg1;
    uint32_t wa[N + 2], wb[N + 2];
    mp_lin2_signed<N>(wa, a, af0, mf0, b, ag0, mg0);
    mp_lin2_signed<N>(wb, a, af1, mf1, b, ag1, mg1);
    const uint32_t na = 0u - (wa[N + 1] >> 31), nb = 0u - (wb[N + 1] >> 31);
    {
      uint64_t c1 = na & 1u, c2 = nb & 1u;
#pragma unroll
      for (int j = 0; j < N + 2; j++) {
        uint64_t z1 = (uint64_t)(wa[j] ^ na) + c1; wa[j] = (uint32_t)z1; c1 = z1 >> 32;
        uint64_t z2 = (uint64_t)(wb[j] ^ nb) + c2; wb[j] = (uint32_t)z2; c2 = z2 >> 32;
      }
    }
#pragma unroll
    for (int j = 0; j < N; j++) { a[j] = (wa[j] >> 30) | (wa[j + 1] << 2); b[j] = (wb[j] >> 30) | (wb[j + 1] << 2); }
    mf0 ^= na; mg0 ^= na; mf1 ^= nb; mg1 ^= nb;   // negated row: every entry changes sign (zero entries: magnitude 0)
    // ---- (u, v) <- the same rows mod p: a negative entry multiplies p - u instead of u
    uint32_t un[N], vn[N];
    un[0] = ptx::sub_cc(C::mod(0), u[0]);
#pragma unroll
    for (int j = 1; j < N; j++) un[j] = ptx::subc_cc(C::mod(j), u[j]);
    vn[0] = ptx::sub_cc(C::mod(0), v[0]);
#pragma unroll
    for (int j = 1; j < N; j++) vn[j] = ptx::subc_cc(C::mod(j), v[j]);
    uint32_t x0[N], y0[N], x1[N], y1[N];
#pragma unroll
    for (int j = 0; j < N; j++) {
      x0[j] = mf0 ? un[j] : u[j]; y0[j] = mg0 ? vn[j] : v[j];
      x1[j] = mf1 ? un[j] : u[j]; y1[j] = mg1 ? vn[j] : v[j];
    }
    mp_lin2_mont30<C>(u, x0, af0, y0, ag0);
    mp_lin2_mont30<C>(v, x1, af1, y1, ag1);
  }
  Fp<C> r, r3;
#pragma unroll
  for (int j = 0; j < N; j++) { r.v[j] = v[j]; r3.v[j] = C::r3(j); }
  fp_mul_c(out, r, r3);
}

}  // namespace b2k
