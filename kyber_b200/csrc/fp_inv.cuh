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

}  // namespace b2k
