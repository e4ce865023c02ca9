// msm_affine.cuh -- affine pair-tree rounds of the MSM's bucket accumulation.
//
// The bucket sums of the Pippenger pipeline (msm.cuh; reference loops share/poly.go:461-473, sign/bdn/bdn.go:126-161)
// are bound by the integer multiply pipe, not by HBM: a mixed XYZZ addition costs 10 field
// products.  An AFFINE addition costs 1 inversion + 3 products, and inversions batch (Montgomery's trick: 3 products per
// element + one shared inversion), i.e. 6 products per addition -- if independent additions are available.  The serial
// "acc += P" of a bucket is not; a pair tree is: one ROUND replaces the k sorted operands of every bucket by the
// ceil(k/2) sums of neighbouring pairs (an odd last operand is carried over), and the rounds' outputs are again affine
// points grouped by bucket.  R rounds shrink the entry list 2^R-fold (trading extra HBM traffic, which the pass has to
// spare, for 40 % fewer products); the XYZZ slices of msm.cuh finish what is left.
//
// One thread owns B consecutive OUTPUT positions of a round:
//   forward : fetch the two operands, classify, d_j = x2 - x1 (2 y1 for a doubling, 1 when nothing is to invert),
//             pre[j] = d_0 ... d_{j-1}                                                    1 product
//   invert  : the running product, once per thread (branch-free binary GCD, fp_inv.cuh)
//   backward: 1/d_j = inv pre[j], inv *= d_j, lambda = (y2 - y1)/d_j, x3 = lambda^2 - x1 - x2,
//             y3 = lambda (x1 - x3) - y1                                                   5 products
// Every exceptional case of the group law is handled (operand at infinity, P + P, P + (-P)): skewed scalar sets put equal
// points into one bucket.
#pragma once
#include "msm.cuh"
#include "fp_inv.cuh"

namespace b2k {

template <class C> B2K_D void f_inv_bg(Fp<C>& r, const Fp<C>& a) { fp_inv_bingcd(r, a); }
template <class C> B2K_D void f_inv_bg(Fp2<C>& r, const Fp2<C>& a) {
  Fp<C> n, t;
  fp_sqr_c(n, a.c0); fp_sqr_c(t, a.c1); fp_add(n, n, t);
  fp_inv_bingcd(n, n);
  fp_mul_c(r.c0, a.c0, n); fp_mul_c(t, a.c1, n); fp_neg(r.c1, t);
}

// (Round 2 tried software prefetch here -- `prefetch.global.L1` of the next output's operands and prefix product one iteration
//  ahead, because ncu shows `long_scoreboard` stalls in the round-1 kernels: measured SLOWER, accumulate 5.48 -> 5.61 ms,
//  profiles/r02_notes.md; the extra address arithmetic and the L1 traffic cost more than the latency they hide.  Removed.)

constexpr int PT_MAXB = 64;     // outputs per thread (length of the prefix-product array in local memory)
enum : int { PT_COPY1 = 0, PT_COPY2 = 1, PT_ADD = 2, PT_DBL = 3, PT_INF = 4 };

// operands of the output at input position a (pairs start at even offsets of the bucket's run [.., end))
template <class CV, bool FIRST>
B2K_D bool pt_fetch(Affine<typename CV::F>& p1, Affine<typename CV::F>& p2, const Affine<typename CV::F>* in,
                    const uint32_t* entries, uint32_t a, uint32_t end) {
  const bool pair = a + 1 < end;
  if (FIRST) {                                        // round 0 gathers by the sorted (index | sign) entries
    uint32_t v = entries[a];
    p1 = in[v & 0x7fffffffu];
    if (v >> 31) f_neg(p1.y, p1.y);
    if (pair) {
      v = entries[a + 1];
      p2 = in[v & 0x7fffffffu];
      if (v >> 31) f_neg(p2.y, p2.y);
    } else p2 = p1;
  } else {
    p1 = in[a];
    p2 = pair ? in[a + 1] : p1;
  }
  return pair;
}

// what the output is and the denominator d its slope needs (1 when none)
template <class F>
B2K_D int pt_classify(F& d, const Affine<F>& p1, const Affine<F>& p2, bool pair) {
  f_set_one(d);
  if (!pair || aff_is_inf(p2)) return PT_COPY1;
  if (aff_is_inf(p1)) return PT_COPY2;
  F dx;
  f_sub(dx, p2.x, p1.x);
  if (!f_is_zero(dx)) { d = dx; return PT_ADD; }
  if (f_eq(p1.y, p2.y) && !f_is_zero(p1.y)) { f_dbl(d, p1.y); return PT_DBL; }
  return PT_INF;                                      // P + (-P), or a 2-torsion point doubled
}

// thread t of a round: outputs [t B, min((t+1) B, offs_out[total]))
template <class CV, bool FIRST>
B2K_D void msm_pairtree_round(uint32_t t, uint32_t B, uint32_t total, const Affine<typename CV::F>* in,
                              const uint32_t* entries, const uint32_t* offs_in, const uint32_t* offs_out,
                              Affine<typename CV::F>* out) {
  using F = typename CV::F;
  const uint32_t nout = offs_out[total];
  const uint32_t q0 = t * B;
  if (q0 >= nout) return;
  const uint32_t q1 = (nout - q0 < B) ? nout : q0 + B;
  F pre[PT_MAXB];
  uint32_t g = msm_find_bucket(offs_out, total, q0);
  uint32_t os = offs_out[g], oe = offs_out[g + 1], is = offs_in[g], ie = offs_in[g + 1];
  F acc;
  f_set_one(acc);
  for (uint32_t q = q0; q < q1; q++) {
    while (q >= oe) { g++; os = oe; oe = offs_out[g + 1]; is = ie; ie = offs_in[g + 1]; }
    Affine<F> p1, p2;
    F d;
    const bool pair = pt_fetch<CV, FIRST>(p1, p2, in, entries, is + 2 * (q - os), ie);
    pt_classify(d, p1, p2, pair);
    pre[q - q0] = acc;
    f_mul(acc, acc, d);
  }
  F inv;
  f_inv_bg(inv, acc);                                 // every d is non-zero, so is their product
  for (uint32_t q = q1; q-- > q0;) {
    while (q < os) { g--; oe = os; os = offs_out[g]; ie = is; is = offs_in[g]; }
    Affine<F> p1, p2, r;
    F d, dinv, lam, tt;
    const bool pair = pt_fetch<CV, FIRST>(p1, p2, in, entries, is + 2 * (q - os), ie);
    const int kind = pt_classify(d, p1, p2, pair);
    f_mul(dinv, inv, pre[q - q0]);
    if (q > q0) f_mul(inv, inv, d);
    if (kind == PT_DBL) { f_sqr(tt, p1.x); f_dbl(lam, tt); f_add(tt, lam, tt); }   // 3 x^2 (rare)
    else f_sub(tt, p2.y, p1.y);
    f_mul(lam, tt, dinv);
    f_sqr(r.x, lam); f_sub(r.x, r.x, p1.x); f_sub(r.x, r.x, p2.x);
    f_sub(tt, p1.x, r.x); f_mul(r.y, lam, tt); f_sub(r.y, r.y, p1.y);
    if (kind == PT_COPY1) r = p1;
    else if (kind == PT_COPY2) r = p2;
    else if (kind == PT_INF) aff_set_inf(r);
    out[q] = r;
  }
}

// ---- the same round as three kernels (forward products / inversions / backward additions) ------------------------------
// One fused kernel keeps three very different code regions (one inlined product; the 2 000-instruction inversion; five
// inlined products) resident at once and runs them all at the register budget of the widest; split, every phase has
// compact code and its own occupancy, the prefix products travel through a coalesced global array pre[j * T + t] and the
// running product / its inverse through accs[t].  The forward pass reads only the x coordinates (dx is all it needs;
// x = 0 or dx = 0 fall back to the full classification).
template <class CV, bool FIRST>
B2K_D void pt_denominator(typename CV::F& d, const Affine<typename CV::F>* in, const uint32_t* entries, uint32_t a, uint32_t end) {
  using F = typename CV::F;
  f_set_one(d);
  if (a + 1 >= end) return;                           // single operand: carried over
  F x1, x2, dx;
  if (FIRST) { x1 = in[entries[a] & 0x7fffffffu].x; x2 = in[entries[a + 1] & 0x7fffffffu].x; }
  else { x1 = in[a].x; x2 = in[a + 1].x; }
  f_sub(dx, x2, x1);
  if (!f_is_zero(x1) && !f_is_zero(x2) && !f_is_zero(dx)) { d = dx; return; }
  Affine<F> p1, p2;                                   // rare: possible infinity / doubling / cancellation
  const bool pair = pt_fetch<CV, FIRST>(p1, p2, in, entries, a, end);
  pt_classify(d, p1, p2, pair);
}

// forward: T = number of threads of the round (stride of pre[]); thread t leaves its running product in accs[t]
template <class CV, bool FIRST>
B2K_D void msm_pairtree_forward(uint32_t t, uint32_t B, uint32_t T, uint32_t total, const Affine<typename CV::F>* in,
                                const uint32_t* entries, const uint32_t* offs_in, const uint32_t* offs_out,
                                typename CV::F* pre, typename CV::F* accs) {
  using F = typename CV::F;
  const uint32_t nout = offs_out[total];
  const uint32_t q0 = t * B;
  if (q0 >= nout) return;
  const uint32_t q1 = (nout - q0 < B) ? nout : q0 + B;
  uint32_t g = msm_find_bucket(offs_out, total, q0);
  uint32_t os = offs_out[g], oe = offs_out[g + 1], is = offs_in[g], ie = offs_in[g + 1];
  F acc;
  f_set_one(acc);
  for (uint32_t q = q0; q < q1; q++) {
    while (q >= oe) { g++; os = oe; oe = offs_out[g + 1]; is = ie; ie = offs_in[g + 1]; }
    F d;
    pt_denominator<CV, FIRST>(d, in, entries, is + 2 * (q - os), ie);
    pre[(size_t)(q - q0) * T + t] = acc;
    f_mul_i(acc, acc, d);
  }
  accs[t] = acc;
}

// inversions: one per thread that has outputs
template <class F>
B2K_D void msm_pairtree_invert(uint32_t t, uint32_t B, uint32_t total, const uint32_t* offs_out, F* accs) {
  if ((uint64_t)t * B >= offs_out[total]) return;
  F v = accs[t], r;
  f_inv_bg(r, v);
  accs[t] = r;
}

// backward: the additions themselves
template <class CV, bool FIRST>
B2K_D void msm_pairtree_backward(uint32_t t, uint32_t B, uint32_t T, uint32_t total, const Affine<typename CV::F>* in,
                                 const uint32_t* entries, const uint32_t* offs_in, const uint32_t* offs_out,
                                 const typename CV::F* pre, const typename CV::F* accs, Affine<typename CV::F>* out) {
  using F = typename CV::F;
  const uint32_t nout = offs_out[total];
  const uint32_t q0 = t * B;
  if (q0 >= nout) return;
  const uint32_t q1 = (nout - q0 < B) ? nout : q0 + B;
  uint32_t g = msm_find_bucket(offs_out, total, q1 - 1);
  uint32_t os = offs_out[g], oe = offs_out[g + 1], is = offs_in[g], ie = offs_in[g + 1];
  (void)oe;
  F inv = accs[t];
  for (uint32_t q = q1; q-- > q0;) {
    while (q < os) { g--; os = offs_out[g]; ie = is; is = offs_in[g]; }
    Affine<F> p1, p2, r;
    F d, dinv, lam, tt;
    const bool pair = pt_fetch<CV, FIRST>(p1, p2, in, entries, is + 2 * (q - os), ie);
    const int kind = pt_classify(d, p1, p2, pair);
    F pj = pre[(size_t)(q - q0) * T + t];
    f_mul_i(dinv, inv, pj);
    if (q > q0) f_mul_i(inv, inv, d);
    if (kind == PT_DBL) { f_sqr(tt, p1.x); f_dbl(lam, tt); f_add(tt, lam, tt); }   // (rare: stays a call in the compact layout)
    else f_sub(tt, p2.y, p1.y);
    f_mul_i(lam, tt, dinv);
    f_sqr_i(r.x, lam); f_sub(r.x, r.x, p1.x); f_sub(r.x, r.x, p2.x);
    f_sub(tt, p1.x, r.x); f_mul_i(r.y, lam, tt); f_sub(r.y, r.y, p1.y);
    if (kind == PT_COPY1) r = p1;
    else if (kind == PT_COPY2) r = p2;
    else if (kind == PT_INF) aff_set_inf(r);
    out[q] = r;
  }
}

}  // namespace b2k
