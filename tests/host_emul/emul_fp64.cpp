// Host emulation of tools/probe/msm_slice_fp64.cuh (test infrastructure; the probe is not part of the product library):
// the MSM pipeline of emul.cpp with every slice, no slice, or alternate groups of slices accumulated on the FP64-form field.
#include <cfenv>
#include <cstring>
#include <vector>
#include "../../kyber_b200/csrc/constants.cuh"
#include "../../kyber_b200/csrc/fp.cuh"
#include "../../tools/probe/msm_slice_fp64.cuh"
#include "../../tools/probe/fpd_overloads.cuh"
#include "../../kyber_b200/csrc/msm_affine.cuh"
using namespace b2k;

// mode: 0 = all slices in the IMAD form, 1 = all in the FP64 form, 2 = alternate pairs of slices (stand-in for warp specialisation)
extern "C" int emul_bls12381_g1_msm_fp64(size_t n, const uint8_t* scalars, const uint8_t* pts, int c, int m, int L, int mode, uint8_t* out) {
  using CV = Bls381G1;
  using F = CV::F;
  const int old = std::fegetround();
  std::fesetround(FE_TOWARDZERO);
  MsmPlan pl; pl.c = c; pl.W = (256 + c - 1) / c; pl.nb = 1 << (c - 1); pl.m = m;
  uint32_t K[9] = {0};
  for (int w = 0; w < pl.W; w++) { int bit = c * w + c - 1; if (bit < 288) K[bit >> 5] |= 1u << (bit & 31); }
  memcpy(pl.K, K, sizeof K);
  std::vector<Affine<F>> P(n);
  for (size_t i = 0; i < n; i++) CV::load(P[i], pts + CV::IN_BYTES * i);
  const size_t total = (size_t)pl.W * pl.nb;
  std::vector<uint32_t> counts(total + 1, 0), offs(total + 1, 0);
  for (size_t i = 0; i < n; i++) {
    Scalar256 s; scalar_load_be(s, scalars + 32 * i);
    if (!scalar_in_range<CV::ScalarField>(s)) { std::fesetround(old); return -3; }
    uint32_t sp[9]; msm_recode(sp, s, pl.K);
    for (int w = 0; w < pl.W; w++) { int d = msm_digit(sp, c, w); if (d) counts[(size_t)w * pl.nb + (d < 0 ? -d : d) - 1]++; }
  }
  for (size_t g = 0; g < total; g++) offs[g + 1] = offs[g] + counts[g];
  std::vector<uint32_t> cursor(offs.begin(), offs.end()), entries(offs[total]);
  for (size_t i = 0; i < n; i++) {
    Scalar256 s; scalar_load_be(s, scalars + 32 * i);
    uint32_t sp[9]; msm_recode(sp, s, pl.K);
    for (int w = 0; w < pl.W; w++) { int d = msm_digit(sp, c, w); if (d) { size_t g = (size_t)w * pl.nb + (d < 0 ? -d : d) - 1; entries[cursor[g]++] = (uint32_t)i | (d < 0 ? 0x80000000u : 0); } }
  }
  std::vector<Xyzz<F>> B(total);
  memset((void*)B.data(), 0, total * sizeof(Xyzz<F>));
  const uint32_t E = offs[total], S = (E + L - 1) / L;
  std::vector<Xyzz<F>> spart(2 * (size_t)S + 2);
  for (uint32_t j = 0; j < S; j++) {
    const bool fp64 = mode == 1 || (mode == 2 && ((j >> 1) & 1));
    if (fp64) msm_accumulate_slice_fp64<CV>(j, (uint32_t)L, (uint32_t)total, P.data(), offs.data(), entries.data(), B.data(), spart.data());
    else msm_accumulate_slice<CV>(j, (uint32_t)L, (uint32_t)total, P.data(), offs.data(), entries.data(), B.data(), spart.data());
  }
  for (size_t g = 0; g < total; g++) {
    if (msm_fixup_bucket<CV, 3>((uint32_t)g, (uint32_t)L, offs.data(), B.data(), spart.data())) {
      uint32_t s0 = offs[g], t0 = offs[g + 1], j0 = s0 / L, j1 = (t0 - 1) / L;
      Xyzz<F> a = spart[2 * (size_t)j0 + 1];
      for (uint32_t j = j0 + 1; j <= j1; j++) xyzz_add(a, a, spart[2 * (size_t)j]);
      B[g] = a;
    }
  }
  const int T = pl.nb / m;
  std::vector<Xyzz<F>> wsum(pl.W);
  for (int w = 0; w < pl.W; w++) {
    Xyzz<F> a; xyzz_set_inf(a);
    for (int t = 0; t < T; t++) { Xyzz<F> part; msm_reduce_chunk<CV>(part, &B[(size_t)w * pl.nb], t, m); xyzz_add(a, a, part); }
    wsum[w] = a;
  }
  Xyzz<F> r; msm_horner<CV>(r, wsum.data(), pl.W, c);
  Affine<F> a; xyzz_to_affine(a, r);
  CV::store(out, a);
  std::fesetround(old);
  return 0;
}


// The affine pair-tree rounds (msm_affine.cuh: forward / invert / backward, the three-kernel form) instantiated on the FP64-form
// field: operands converted to Affine<FpD> once, `rounds` rounds with PB outputs per thread, outputs converted back and finished
// by the library's XYZZ slices (DIRECT form), fix-up, reduction and Horner.
extern "C" int emul_bls12381_g1_msm_rounds_fp64(size_t n, const uint8_t* scalars, const uint8_t* pts, int c, int m, int L, int rounds, int PB,
                                                uint8_t* out) {
  using CV = Bls381G1;
  using CVD = Bls381G1D;
  using F = CV::F;
  const int old = std::fegetround();
  std::fesetround(FE_TOWARDZERO);
  MsmPlan pl; pl.c = c; pl.W = (256 + c - 1) / c; pl.nb = 1 << (c - 1); pl.m = m;
  uint32_t K[9] = {0};
  for (int w = 0; w < pl.W; w++) { int bit = c * w + c - 1; if (bit < 288) K[bit >> 5] |= 1u << (bit & 31); }
  memcpy(pl.K, K, sizeof K);
  std::vector<Affine<F>> P(n);
  std::vector<Affine<FpD>> Pd(n);
  for (size_t i = 0; i < n; i++) { CV::load(P[i], pts + CV::IN_BYTES * i); FpdConv<Affine<FpD>>::load(Pd[i], P[i]); }
  const size_t total = (size_t)pl.W * pl.nb;
  std::vector<uint32_t> counts(total + 1, 0), offs(total + 1, 0);
  for (size_t i = 0; i < n; i++) {
    Scalar256 s; scalar_load_be(s, scalars + 32 * i);
    uint32_t sp[9]; msm_recode(sp, s, pl.K);
    for (int w = 0; w < pl.W; w++) { int d = msm_digit(sp, c, w); if (d) counts[(size_t)w * pl.nb + (d < 0 ? -d : d) - 1]++; }
  }
  for (size_t g = 0; g < total; g++) offs[g + 1] = offs[g] + counts[g];
  std::vector<uint32_t> cursor(offs.begin(), offs.end()), entries(offs[total]);
  for (size_t i = 0; i < n; i++) {
    Scalar256 s; scalar_load_be(s, scalars + 32 * i);
    uint32_t sp[9]; msm_recode(sp, s, pl.K);
    for (int w = 0; w < pl.W; w++) { int d = msm_digit(sp, c, w); if (d) { size_t g = (size_t)w * pl.nb + (d < 0 ? -d : d) - 1; entries[cursor[g]++] = (uint32_t)i | (d < 0 ? 0x80000000u : 0); } }
  }
  std::vector<Affine<FpD>> cur;
  for (int r = 0; r < rounds; r++) {
    std::vector<uint32_t> no(total + 1, 0);
    for (size_t g = 0; g < total; g++) no[g + 1] = no[g] + ((offs[g + 1] - offs[g] + 1) >> 1);
    std::vector<Affine<FpD>> nxt(no[total] + 1);
    const uint32_t T = (no[total] + PB - 1) / PB + 1;
    std::vector<FpD> pre((size_t)T * PB), accs(T);
    for (uint32_t t = 0; t < T; t++) {
      if (r == 0) msm_pairtree_forward<CVD, true>(t, (uint32_t)PB, T, (uint32_t)total, Pd.data(), entries.data(), offs.data(), no.data(), pre.data(), accs.data());
      else msm_pairtree_forward<CVD, false>(t, (uint32_t)PB, T, (uint32_t)total, cur.data(), nullptr, offs.data(), no.data(), pre.data(), accs.data());
    }
    for (uint32_t t = 0; t < T; t++) msm_pairtree_invert<FpD>(t, (uint32_t)PB, (uint32_t)total, no.data(), accs.data());
    for (uint32_t t = 0; t < T; t++) {
      if (r == 0) msm_pairtree_backward<CVD, true>(t, (uint32_t)PB, T, (uint32_t)total, Pd.data(), entries.data(), offs.data(), no.data(), pre.data(), accs.data(), nxt.data());
      else msm_pairtree_backward<CVD, false>(t, (uint32_t)PB, T, (uint32_t)total, cur.data(), nullptr, offs.data(), no.data(), pre.data(), accs.data(), nxt.data());
    }
    cur.swap(nxt);
    offs.swap(no);
  }
  std::vector<Affine<F>> back(cur.size());
  for (size_t i = 0; i < cur.size(); i++) { fpd_store(back[i].x, cur[i].x); fpd_store(back[i].y, cur[i].y); }
  std::vector<Xyzz<F>> B(total);
  memset((void*)B.data(), 0, total * sizeof(Xyzz<F>));
  const uint32_t E = offs[total], S = (E + L - 1) / L;
  std::vector<Xyzz<F>> spart(2 * (size_t)S + 2);
  for (uint32_t j = 0; j < S; j++) msm_accumulate_slice<CV, true>(j, (uint32_t)L, (uint32_t)total, back.data(), offs.data(), nullptr, B.data(), spart.data());
  for (size_t g = 0; g < total; g++) {
    if (msm_fixup_bucket<CV, 3>((uint32_t)g, (uint32_t)L, offs.data(), B.data(), spart.data())) {
      uint32_t s0 = offs[g], t0 = offs[g + 1], j0 = s0 / L, j1 = (t0 - 1) / L;
      Xyzz<F> a = spart[2 * (size_t)j0 + 1];
      for (uint32_t j = j0 + 1; j <= j1; j++) xyzz_add(a, a, spart[2 * (size_t)j]);
      B[g] = a;
    }
  }
  const int T = pl.nb / m;
  std::vector<Xyzz<F>> wsum(pl.W);
  for (int w = 0; w < pl.W; w++) {
    Xyzz<F> a; xyzz_set_inf(a);
    for (int t = 0; t < T; t++) { Xyzz<F> part; msm_reduce_chunk<CV>(part, &B[(size_t)w * pl.nb], t, m); xyzz_add(a, a, part); }
    wsum[w] = a;
  }
  Xyzz<F> r; msm_horner<CV>(r, wsum.data(), pl.W, c);
  Affine<F> a; xyzz_to_affine(a, r);
  CV::store(out, a);
  std::fesetround(old);
  return 0;
}
