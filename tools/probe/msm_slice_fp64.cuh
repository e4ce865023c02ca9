// msm_slice_fp64.cuh -- PROBE (next round's dual-pipe accumulate pass; not instantiated by the product library): the balanced-slice
// body of the MSM (kyber_b200/csrc/msm.cuh: msm_accumulate_slice) with the bucket accumulator and the additions on the FP64-pipe
// field (fp_dfma.cuh / ec_dfma.cuh, Montgomery radix 2^384 = the library's).  Same inputs, same outputs, same control flow: a warp
// of the accumulate kernel can run either body, which is what lets the kernel split its warps between the FMA-heavy pipe
// (IMAD form) and the FP64 pipe.  Operands and results cross between the forms by re-packing bits (from_u32 / to_u32).
#pragma once
#include "../../kyber_b200/csrc/msm.cuh"
#include "ec_dfma.cuh"

namespace b2k {

template <class F>
B2K_D void fp64_store_xyzz(Xyzz<F>& r, const dfma::Xyzz& a) {
  if (dfma::fp_is_zero(a.ZZ)) { xyzz_set_inf(r); return; }
  dfma::to_u32(r.X.v, a.X); dfma::to_u32(r.Y.v, a.Y); dfma::to_u32(r.ZZ.v, a.ZZ); dfma::to_u32(r.ZZZ.v, a.ZZZ);
}
template <class F>
B2K_D void fp64_load_xyzz(dfma::Xyzz& r, const Xyzz<F>& a) {
  dfma::from_u32(r.X, a.X.v); dfma::from_u32(r.Y, a.Y.v); dfma::from_u32(r.ZZ, a.ZZ.v); dfma::from_u32(r.ZZZ, a.ZZZ.v);
}
B2K_D void fp64_set_inf(dfma::Xyzz& a) {
#pragma unroll
  for (int i = 0; i < dfma::L; i++) { a.X.v[i] = 0.0; a.Y.v[i] = 0.0; a.ZZ.v[i] = 0.0; a.ZZZ.v[i] = 0.0; }
}

// acc += +-q with q in the library's affine form
template <class F>
B2K_D void fp64_madd(dfma::Xyzz& acc, const Affine<F>& q, bool negate, const dfma::Fp& one) {
  if (aff_is_inf(q)) return;
  F qy;
  if (negate) f_neg(qy, q.y); else qy = q.y;
  dfma::Affine d;
  dfma::from_u32(d.x, q.x.v);
  dfma::from_u32(d.y, qy.v);
  if (dfma::xyzz_madd<true>(acc, d, one) == 2) {        // same point: the rare doubling goes through the library form
    Xyzz<F> t, t2;
    t.X = q.x; t.Y = qy; f_set_one(t.ZZ); f_set_one(t.ZZZ);
    xyzz_dbl(t2, t);
    fp64_load_xyzz(acc, t2);
  }
}

template <class CV, bool DIRECT = false>
B2K_D void msm_accumulate_slice_fp64(uint32_t j, uint32_t L, uint32_t total, const Affine<typename CV::F>* pts,
                                     const uint32_t* offs, const uint32_t* entries,
                                     Xyzz<typename CV::F>* buckets, Xyzz<typename CV::F>* spart) {
  using F = typename CV::F;
  static_assert(F::N == 12, "the FP64-form field is built for the 381-bit base field of BLS12-381");
  const uint32_t E = offs[total];
  const uint32_t b = j * L;
  if (b >= E) return;
  const uint32_t e = (E - b < L) ? E : b + L;
  uint32_t g = msm_find_bucket(offs, total, b);
  uint32_t gs = offs[g], ge = offs[g + 1];
  dfma::Fp one;
  { F o; f_set_one(o); dfma::from_u32(one, o.v); }
  dfma::Xyzz acc;
  fp64_set_inf(acc);
  Xyzz<F> out;
  for (uint32_t pos = b; pos < e; pos++) {
    if (pos == ge) {                       // crossed into a later bucket
      fp64_store_xyzz(out, acc);
      msm_slice_flush<CV>(out, g, gs, ge, j, b, e, buckets, spart);
      fp64_set_inf(acc);
      do { g++; gs = ge; ge = offs[g + 1]; } while (ge <= pos);
    }
    if (DIRECT) {
      fp64_madd(acc, pts[pos], false, one);
    } else {
      const uint32_t v = entries[pos];
      fp64_madd(acc, pts[v & 0x7fffffffu], (v >> 31) != 0, one);
    }
  }
  fp64_store_xyzz(out, acc);
  msm_slice_flush<CV>(out, g, gs, ge, j, b, e, buckets, spart);
}

}  // namespace b2k
