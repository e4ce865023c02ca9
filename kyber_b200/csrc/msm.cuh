// msm.cuh -- per-thread bodies of the batched scalar-multiplication and Pippenger MSM kernels.
//
// Replaces (reference call sites; the reference has no batched entry point, only Go loops):
//   kilic.G1Elt.Mul                    pairing/bls12381/kilic/g1.go:110-116   (one s*P per call)
//   bn254 curvePoint.Mul               pairing/bn254/curve.go:196-218
//   the loops  Tmp.Mul(l_i, y_i); Acc.Add(Acc, Tmp)   share/poly.go:461-473 (RecoverCommit)
//              sig.Mul(coef, sig); agg.Add(...)       sign/bdn/bdn.go:126-161 (AggregateSignatures)
//
// MSM pipeline (window c bits, W = ceil(256/c) windows, signed digits -> 2^(c-1) buckets/window):
//   digits+count  : s' = s + K (K = sum of 2^(c-1) at every window) so that window digits of s' minus
//                   2^(c-1) are the signed digits; histogram of (window, |digit|) by atomics
//   scan          : exclusive prefix sum of the histogram -> bucket start offsets
//   scatter       : counting sort of (point index | sign) by bucket
//   accumulate    : one thread per bucket, XYZZ += +-P (mixed add 8M+2S), points gathered by index
//   reduce        : per (window, chunk of m buckets): running sums  A = sum (k-k0) B_k,  S = sum B_k,
//                   partial = A + k0*S ; then per-window tree sum of the partials
//   final         : Horner over windows (c doublings per window), to affine, to wire bytes
// The per-thread bodies below are plain functions of explicit indices so that tests/host_emul can run
// them on the CPU; the __global__ wrappers live in kernels.cu.
#pragma once
#include <stddef.h>
#include "curves.cuh"

namespace b2k {

struct MsmPlan {
  int c;           // window bits, 2..16
  int W;           // windows
  int nb;          // buckets per window = 2^(c-1)
  int m;           // reduction chunk (power of two, divides nb)
  uint32_t K[9];   // recoding offset, little-endian limbs
};

// signed digit of window w of s' (s' given as 9 limbs):  returns d in [-2^(c-1), 2^(c-1)-1]
B2K_D int msm_digit(const uint32_t* sp, int c, int w) {
  int bit = c * w;
  int limb = bit >> 5, sh = bit & 31;
  uint64_t two = (uint64_t)sp[limb] | ((uint64_t)(limb + 1 < 9 ? sp[limb + 1] : 0u) << 32);
  uint32_t e = (uint32_t)(two >> sh) & ((1u << c) - 1u);
  return (int)e - (1 << (c - 1));
}

// s' = s + K  (9 limbs)
B2K_D void msm_recode(uint32_t* sp, const Scalar256& s, const uint32_t* K) {
  sp[0] = ptx::add_cc(s.v[0], K[0]);
#pragma unroll
  for (int j = 1; j < 8; j++) sp[j] = ptx::addc_cc(s.v[j], K[j]);
  sp[8] = ptx::addc(0, K[8]);
}

// ---- accumulate: one bucket -----------------------------------------------------------------
// entries[start..end) hold (point index | sign << 31)
template <class CV>
B2K_D void msm_accumulate_bucket(Xyzz<typename CV::F>& acc, const Affine<typename CV::F>* pts,
                                 const uint32_t* entries, uint32_t start, uint32_t end) {
  xyzz_set_inf(acc);
  for (uint32_t e = start; e < end; e++) {
    uint32_t v = entries[e];
    Affine<typename CV::F> q = pts[v & 0x7fffffffu];
    xyzz_madd(acc, acc, q, (v >> 31) != 0);
  }
}

// ---- reduce: one chunk of m consecutive buckets ------------------------------------------------
// buckets of this window: B[0..nb) where B[k] holds digit magnitude k+1.  Chunk t covers k in
// [t*m, (t+1)*m).  partial = sum_{k} (k+1) * B[k].
template <class CV>
B2K_D void msm_reduce_chunk(Xyzz<typename CV::F>& out, const Xyzz<typename CV::F>* B, int t, int m) {
  using X = Xyzz<typename CV::F>;
  X run, acc;
  xyzz_set_inf(run);
  xyzz_set_inf(acc);
  for (int k = m - 1; k >= 0; k--) {
    X b = B[t * m + k];
    xyzz_add(run, run, b);
    xyzz_add(acc, acc, run);      // after the loop: acc = sum (k+1) B[t*m+k], run = sum B
  }
  if (t != 0) {
    X off;
    xyzz_mul_small(off, run, (uint32_t)(t * m));
    xyzz_add(acc, acc, off);
  }
  out = acc;
}

// ---- final: Horner over window sums --------------------------------------------------------------
template <class CV>
B2K_D void msm_horner(Xyzz<typename CV::F>& out, const Xyzz<typename CV::F>* wsum, int W, int c) {
  using X = Xyzz<typename CV::F>;
  X acc = wsum[W - 1];
  for (int w = W - 2; w >= 0; w--) {
    for (int i = 0; i < c; i++) xyzz_dbl(acc, acc);
    X t = wsum[w];
    xyzz_add(acc, acc, t);
  }
  out = acc;
}

// ---- independent scalar multiplication -----------------------------------------------------------
// k*P, MSB-first double-and-add over the 256-bit scalar (Jacobian, mixed additions).
template <class CV>
B2K_D void scalar_mul(Jac<typename CV::F>& r, const Scalar256& k, const Affine<typename CV::F>& p) {
  Jac<typename CV::F> acc;
  jac_set_inf(acc);
  bool started = false;
  for (int i = 255; i >= 0; i--) {
    if (started) jac_dbl(acc, acc);
    if ((k.v[i >> 5] >> (i & 31)) & 1u) {
      jac_madd(acc, acc, p);
      started = true;
    }
  }
  r = acc;
}

}  // namespace b2k

// =================================================================================================
// v2 accumulate: fixed-length slices of the sorted entry list (perfectly balanced, skew-proof)
// =================================================================================================
// The sorted entries are cut into slices of L consecutive entries regardless of bucket boundaries;
// one thread per slice performs exactly L mixed additions.  A bucket that lies entirely inside a
// slice is written straight to buckets[]; a bucket cut by a slice boundary gets one partial sum per
// slice it touches (spart[2*j+0] if it started before slice j, spart[2*j+1] if it starts inside slice
// j and runs past its end) and msm_fixup_bucket() adds the partials.  Work per thread is independent
// of the scalar distribution (all-equal scalars, 128-bit BDN coefficients, short top windows ...).
namespace b2k {

// largest g in [0,total) with offs[g] <= pos   (offs non-decreasing, offs[0] = 0, offs[total] > pos)
B2K_D uint32_t msm_find_bucket(const uint32_t* offs, uint32_t total, uint32_t pos) {
  uint32_t lo = 0, hi = total;
  while (hi - lo > 1) {
    uint32_t mid = lo + ((hi - lo) >> 1);
    if (offs[mid] <= pos) lo = mid; else hi = mid;
  }
  return lo;
}

template <class CV>
B2K_D void msm_slice_flush(const Xyzz<typename CV::F>& acc, uint32_t g, uint32_t s, uint32_t t, uint32_t j,
                           uint32_t b, uint32_t e, Xyzz<typename CV::F>* buckets, Xyzz<typename CV::F>* spart) {
  if (s >= b && t <= e) buckets[g] = acc;
  else spart[2 * (size_t)j + (s < b ? 0 : 1)] = acc;
}

template <class CV>
B2K_D void msm_accumulate_slice(uint32_t j, uint32_t L, uint32_t total, const Affine<typename CV::F>* pts,
                                const uint32_t* offs, const uint32_t* entries,
                                Xyzz<typename CV::F>* buckets, Xyzz<typename CV::F>* spart) {
  const uint32_t E = offs[total];
  const uint32_t b = j * L;
  if (b >= E) return;
  const uint32_t e = (E - b < L) ? E : b + L;
  uint32_t g = msm_find_bucket(offs, total, b);
  uint32_t gs = offs[g], ge = offs[g + 1];
  Xyzz<typename CV::F> acc;
  xyzz_set_inf(acc);
  for (uint32_t pos = b; pos < e; pos++) {
    if (pos == ge) {                       // crossed into a later bucket
      msm_slice_flush<CV>(acc, g, gs, ge, j, b, e, buckets, spart);
      xyzz_set_inf(acc);
      do { g++; gs = ge; ge = offs[g + 1]; } while (ge <= pos);
    }
    uint32_t v = entries[pos];
    Affine<typename CV::F> q = pts[v & 0x7fffffffu];
    xyzz_madd(acc, acc, q, (v >> 31) != 0);
  }
  msm_slice_flush<CV>(acc, g, gs, ge, j, b, e, buckets, spart);
}

// Combine the partials of a bucket that was cut by slice boundaries.  Returns true when the bucket has
// more than SERIAL_MAX partials (left to the block-parallel path), false when done / nothing to do.
template <class CV, int SERIAL_MAX>
B2K_D bool msm_fixup_bucket(uint32_t g, uint32_t L, const uint32_t* offs, Xyzz<typename CV::F>* buckets,
                            const Xyzz<typename CV::F>* spart) {
  uint32_t s = offs[g], t = offs[g + 1];
  if (t == s) return false;
  uint32_t j0 = s / L, j1 = (t - 1) / L;
  if (j0 == j1) return false;
  if (j1 - j0 + 1 > (uint32_t)SERIAL_MAX) return true;
  Xyzz<typename CV::F> acc = spart[2 * (size_t)j0 + 1];
  for (uint32_t j = j0 + 1; j <= j1; j++) {
    Xyzz<typename CV::F> p = spart[2 * (size_t)j];
    xyzz_add(acc, acc, p);
  }
  buckets[g] = acc;
  return false;
}

}  // namespace b2k
