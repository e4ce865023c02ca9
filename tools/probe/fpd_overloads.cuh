// fpd_overloads.cuh -- PROBE: the FP64-form field (fp_dfma.cuh, Montgomery radix 2^384) behind the library's generic field
// interface f_add / f_sub / f_mul / f_sqr / ... (kyber_b200/csrc/tower.cuh), so that the EC templates of ec.cuh (Affine<F>,
// Xyzz<F>, xyzz_madd, xyzz_add, xyzz_dbl ...) instantiate on it unchanged: Xyzz<FpD> is a bucket on the FP64 pipe.
// Next round: separate storage type (12 x u32, as now) from compute type in msm.cuh / msm_affine.cuh and give half of the warps FpD.
#pragma once
#include "../../kyber_b200/csrc/tower.cuh"
#include "../../kyber_b200/csrc/ec.cuh"
#include "../../kyber_b200/csrc/fp_inv.cuh"
#include "ec_dfma.cuh"

namespace b2k {

using FpD = dfma::Fp;

B2K_D void fpd_load(FpD& r, const Fp<Bls381Fp>& a) { dfma::from_u32(r, a.v); }
B2K_D void fpd_store(Fp<Bls381Fp>& r, const FpD& a) { dfma::to_u32(r.v, a); }

}  // namespace b2k

// The overloads live in namespace dfma so that argument-dependent lookup finds them from inside the library's templates.
namespace dfma {
using b2k::FpD;
using b2k::fpd_load;
using b2k::fpd_store;
B2K_D void f_add(FpD& r, const FpD& a, const FpD& b) { dfma::fp_add(r, a, b); }
B2K_D void f_sub(FpD& r, const FpD& a, const FpD& b) { dfma::fp_sub(r, a, b); }
B2K_D void f_mul(FpD& r, const FpD& a, const FpD& b) { dfma::mont_mul384(r, a, b); }
B2K_D void f_sqr(FpD& r, const FpD& a) { dfma::mont_mul384(r, a, a); }
B2K_D void f_mul_i(FpD& r, const FpD& a, const FpD& b) { dfma::mont_mul384(r, a, b); }   // the pair-tree rounds' inlined spelling
B2K_D void f_sqr_i(FpD& r, const FpD& a) { dfma::mont_mul384(r, a, a); }
B2K_D void f_set_zero(FpD& r) {
#pragma unroll
  for (int i = 0; i < dfma::L; i++) r.v[i] = 0.0;
}
B2K_D void f_neg(FpD& r, const FpD& a) { FpD z; f_set_zero(z); dfma::fp_sub(r, z, a); }
B2K_D void f_dbl(FpD& r, const FpD& a) { dfma::fp_add(r, a, a); }
B2K_D bool f_is_zero(const FpD& a) { return dfma::fp_is_zero(a); }
B2K_D bool f_eq(const FpD& a, const FpD& b) {
  bool e = true;
#pragma unroll
  for (int i = 0; i < dfma::L; i++) e = e && (a.v[i] == b.v[i]);     // canonical limbs: equal values have equal limbs
  return e;
}
B2K_D void f_set_one(FpD& r) { b2k::Fp<b2k::Bls381Fp> o; b2k::fp_set_one(o); fpd_load(r, o); }
B2K_D void f_inv_bg(FpD& r, const FpD& a) { b2k::Fp<b2k::Bls381Fp> t, u; fpd_store(t, a); b2k::fp_inv_bingcd(u, t); fpd_load(r, u); }
B2K_D void f_inv(FpD& r, const FpD& a) { b2k::Fp<b2k::Bls381Fp> t, u; fpd_store(t, a); b2k::fp_inv(u, t); fpd_load(r, u); }

}  // namespace dfma

namespace b2k {

// the curve tag the MSM templates (msm_affine.cuh: pair-tree rounds) need: only the field type
struct Bls381G1D { using F = FpD; };

template <class T> struct FpdConv;
template <> struct FpdConv<Affine<FpD>> {
  static B2K_D void load(Affine<FpD>& r, const Affine<Fp<Bls381Fp>>& a) { fpd_load(r.x, a.x); fpd_load(r.y, a.y); }
};
template <> struct FpdConv<Xyzz<FpD>> {
  static B2K_D void load(Xyzz<FpD>& r, const Xyzz<Fp<Bls381Fp>>& a) { fpd_load(r.X, a.X); fpd_load(r.Y, a.Y); fpd_load(r.ZZ, a.ZZ); fpd_load(r.ZZZ, a.ZZZ); }
  static B2K_D void store(Xyzz<Fp<Bls381Fp>>& r, const Xyzz<FpD>& a) { fpd_store(r.X, a.X); fpd_store(r.Y, a.Y); fpd_store(r.ZZ, a.ZZ); fpd_store(r.ZZZ, a.ZZZ); }
};

}  // namespace b2k
