// Host emulation check of tools/probe/fpd_overloads.cuh: the library's own EC templates (ec.cuh) instantiated on the FP64-form
// field must agree, value for value, with their instantiation on the 12 x 32-bit field.
#include <cfenv>
#include <cstring>
#include "../../kyber_b200/csrc/constants.cuh"
#include "../../kyber_b200/csrc/fp.cuh"
#include "../../tools/probe/fpd_overloads.cuh"
#include "../../tools/probe/msm_slice_fp64.cuh"
using namespace b2k;
using F = Fp<Bls381Fp>;

static bool same(const Xyzz<F>& a, const Xyzz<F>& b) {
  if (xyzz_is_inf(a) || xyzz_is_inf(b)) return xyzz_is_inf(a) && xyzz_is_inf(b);
  return fp_eq(a.X, b.X) && fp_eq(a.Y, b.Y) && fp_eq(a.ZZ, b.ZZ) && fp_eq(a.ZZZ, b.ZZZ);
}

// in: acc = 4 x 12 u32 (library Xyzz, Montgomery), q = 2 x 12 u32 (library Affine), r = 4 x 12 u32 (a second Xyzz).
// Runs xyzz_madd (both signs), xyzz_add and xyzz_dbl in both fields; returns a bit mask of disagreements (0 = all equal).
extern "C" int emul_fpd_ec_agree(const uint32_t* acc, const uint32_t* q, const uint32_t* r2) {
  const int old = std::fegetround();
  std::fesetround(FE_TOWARDZERO);
  Xyzz<F> a, b; Affine<F> p;
  memcpy(a.X.v, acc, 48); memcpy(a.Y.v, acc + 12, 48); memcpy(a.ZZ.v, acc + 24, 48); memcpy(a.ZZZ.v, acc + 36, 48);
  memcpy(b.X.v, r2, 48); memcpy(b.Y.v, r2 + 12, 48); memcpy(b.ZZ.v, r2 + 24, 48); memcpy(b.ZZZ.v, r2 + 36, 48);
  memcpy(p.x.v, q, 48); memcpy(p.y.v, q + 12, 48);
  Xyzz<FpD> ad, bd; Affine<FpD> pd;
  FpdConv<Xyzz<FpD>>::load(ad, a); FpdConv<Xyzz<FpD>>::load(bd, b); FpdConv<Affine<FpD>>::load(pd, p);
  int bad = 0;
  for (int neg = 0; neg < 2; neg++) {
    Xyzz<F> w; Xyzz<FpD> wd; Xyzz<F> back;
    xyzz_madd(w, a, p, neg != 0);
    xyzz_madd(wd, ad, pd, neg != 0);
    FpdConv<Xyzz<FpD>>::store(back, wd);
    if (!same(w, back)) bad |= 1 << neg;
  }
  { Xyzz<F> w; Xyzz<FpD> wd; Xyzz<F> back; xyzz_add(w, a, b); xyzz_add(wd, ad, bd); FpdConv<Xyzz<FpD>>::store(back, wd); if (!same(w, back)) bad |= 4; }
  { Xyzz<F> w; Xyzz<FpD> wd; Xyzz<F> back; xyzz_dbl(w, a); xyzz_dbl(wd, ad); FpdConv<Xyzz<FpD>>::store(back, wd); if (!same(w, back)) bad |= 8; }
  for (int neg = 0; neg < 2; neg++) {      // the hand-ordered FP64-form addition of ec_dfma.cuh (what dual_pipe_probe.cu compares by memcmp)
    Xyzz<F> w, back;
    xyzz_madd(w, a, p, neg != 0);
    dfma::Xyzz acc;
    if (xyzz_is_inf(a)) fp64_set_inf(acc); else fp64_load_xyzz(acc, a);
    dfma::Fp one; { F o; f_set_one(o); dfma::from_u32(one, o.v); }
    fp64_madd(acc, p, neg != 0, one);
    fp64_store_xyzz(back, acc);
    if (!same(w, back)) bad |= 32 << neg;
  }
  { F i1, i2; FpD id; fp_inv(i1, a.ZZ); f_inv(id, ad.ZZ); fpd_store(i2, id); if (!fp_eq(i1, i2)) bad |= 16; }
  std::fesetround(old);
  return bad;
}
