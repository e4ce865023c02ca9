// pairing_kernels.cuh -- __global__ wrappers for batched pairings (one pairing / check per thread).
#pragma once
#include "pairing.cuh"
#include "curves.cuh"
#include "codec.cuh"
#include "kernels.cuh"
#include "b2k_ctx.h"

namespace b2k {

// gt[i] = e(g1[i], g2[i])           replaces n x Suite.Pair (kilic/suite.go:70-75)
template <int BLOCK, int MINB>
static __global__ void __launch_bounds__(BLOCK, MINB) k_bls_pair(size_t n, const uint8_t* __restrict__ g1, const uint8_t* __restrict__ g2,
                                                  uint8_t* __restrict__ gt, uint32_t* flags) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<BFp> P;
  Affine<BFp2> Q;
  bool good = load_checked<Bls381G1>(P, g1 + 96 * i);
  good = load_checked<Bls381G2>(Q, g2 + 192 * i) && good;       // a malformed operand counts as infinity: e = 1, FLAG_POINT
  if (!good) atomicOr(flags, FLAG_POINT);
  BFp12 f, e;
  miller_loop<1>(f, &P, &Q);
  final_exponentiation(e, f);
  gt_store(gt + 576 * i, e);
}

// ok[i] = ( e(a1[i], a2[i]) == e(b1[i], b2[i]) )     replaces n x Suite.ValidatePairing
// (kilic/suite.go:57-68): one 2-pair Miller loop (second pair negated) + one final exponentiation.
template <int BLOCK, int MINB>
static __global__ void __launch_bounds__(BLOCK, MINB) k_bls_pairing_check(size_t n, const uint8_t* __restrict__ a1,
                                                           const uint8_t* __restrict__ a2,
                                                           const uint8_t* __restrict__ b1,
                                                           const uint8_t* __restrict__ b2, uint8_t* __restrict__ ok,
                                                           int b2_broadcast, const uint8_t* __restrict__ pre_ok, uint32_t* flags) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (pre_ok && !pre_ok[i]) { ok[i] = 0; return; }       // an operand failed UnmarshalBinary upstream
  Affine<BFp> P[2];
  Affine<BFp2> Q[2];
  bool good = load_checked<Bls381G1>(P[0], a1 + 96 * i);
  good = load_checked<Bls381G2>(Q[0], a2 + 192 * i) && good;
  good = load_checked<Bls381G1>(P[1], b1 + ((b2_broadcast & 2) ? 0 : 96 * i)) && good;      // bit 1: b1 is one shared operand
  good = load_checked<Bls381G2>(Q[1], b2 + ((b2_broadcast & 1) ? 0 : 192 * i)) && good;     // bit 0: b2 is one shared operand
  if (!good) { atomicOr(flags, FLAG_POINT); ok[i] = 0; return; }   // malformed operand (off the curve / coordinate >= p): the check fails
  fp_neg(P[1].y, P[1].y);
  BFp12 f, e;
  miller_loop<2>(f, P, Q);
  final_exponentiation(e, f);
  ok[i] = fp12_is_one(e) ? 1 : 0;
}


// ---- the same two entry points as TWO kernels: Miller loop(s) -> f in global memory -> final exponentiation --------------------
// Why: fused, the kernel needs 255 registers for its field products to be compiled well (tools/codegen_check.py), i.e. 8 warps per
// SM and 65 536 checks = 1.73 waves.  Each half ALONE compiles to clean products at 128 registers (the interprocedural register
// allocation has half the call graph to serve), i.e. 16 warps per SM: all 2 048 warps of 65 536 checks resident at once.
// Measured (profiles/r02n_pairing_split.txt, 65 536 checks): both halves at 128 registers 79.9 ms -- SLOWER than the fused kernel
// (74.3-75.4 ms): with twice the threads the 3-8 KB stacks fall out of L1 and the multiply pipe was already 72 % busy; both halves at
// 255 registers 73.3 ms.  A 2 % gain does not pay for a scratch buffer: the fused kernel stays the default, b2k_set_pairing_variant
// (16 + 2 m + f) selects the split form (m, f = 0: 128 registers, 1: 255).  Cost: 576 bytes per element written and read once.
// status[i]: 2 = f[i] is pending, 0 / 1 = decided.
#ifdef B2K_COMPACT_FIELD
template <int BLOCK, int MINB>
static __global__ void __launch_bounds__(BLOCK, MINB) k_bls_check_miller(size_t n, const uint8_t* __restrict__ a1, const uint8_t* __restrict__ a2,
                                                                         const uint8_t* __restrict__ b1, const uint8_t* __restrict__ b2,
                                                                         BFp12* __restrict__ f_out, uint8_t* __restrict__ ok, int b2_broadcast,
                                                                         const uint8_t* __restrict__ pre_ok, uint32_t* flags) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (pre_ok && !pre_ok[i]) { ok[i] = 0; return; }
  Affine<BFp> P[2];
  Affine<BFp2> Q[2];
  bool good = load_checked<Bls381G1>(P[0], a1 + 96 * i);
  good = load_checked<Bls381G2>(Q[0], a2 + 192 * i) && good;
  good = load_checked<Bls381G1>(P[1], b1 + ((b2_broadcast & 2) ? 0 : 96 * i)) && good;
  good = load_checked<Bls381G2>(Q[1], b2 + ((b2_broadcast & 1) ? 0 : 192 * i)) && good;
  if (!good) { atomicOr(flags, FLAG_POINT); ok[i] = 0; return; }
  fp_neg(P[1].y, P[1].y);
  BFp12 f;
  miller_loop<2>(f, P, Q);
  f_out[i] = f;
  ok[i] = 2;
}
template <int BLOCK, int MINB>
static __global__ void __launch_bounds__(BLOCK, MINB) k_bls_check_final(size_t n, const BFp12* __restrict__ f_in, uint8_t* __restrict__ ok) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || ok[i] != 2) return;
  BFp12 f = f_in[i], e;
  final_exponentiation(e, f);
  ok[i] = fp12_is_one(e) ? 1 : 0;
}
template <int BLOCK, int MINB>
static __global__ void __launch_bounds__(BLOCK, MINB) k_bls_pair_miller(size_t n, const uint8_t* __restrict__ g1, const uint8_t* __restrict__ g2,
                                                                        BFp12* __restrict__ f_out, uint32_t* flags) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<BFp> P;
  Affine<BFp2> Q;
  bool good = load_checked<Bls381G1>(P, g1 + 96 * i);
  good = load_checked<Bls381G2>(Q, g2 + 192 * i) && good;
  if (!good) atomicOr(flags, FLAG_POINT);
  BFp12 f;
  miller_loop<1>(f, &P, &Q);
  f_out[i] = f;
}
template <int BLOCK, int MINB>
static __global__ void __launch_bounds__(BLOCK, MINB) k_bls_pair_final(size_t n, const BFp12* __restrict__ f_in, uint8_t* __restrict__ gt) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  BFp12 f = f_in[i], e;
  final_exponentiation(e, f);
  gt_store(gt + 576 * i, e);
}
// split configuration = 2 * (shape of the Miller kernel) + (shape of the final-exponentiation kernel)
#define B2K_SPLIT_MILLER(X) X(0, 64, 8) X(1, 64, 4)
#define B2K_SPLIT_FINAL(X) X(0, 64, 8) X(1, 64, 4)
inline void launch_pairing_check_split(const b2k_ctx* ctx, int cfg, size_t n, const uint8_t* a1, const uint8_t* a2, const uint8_t* b1,
                                       const uint8_t* b2, uint8_t* ok, int b2_broadcast, const uint8_t* pre_ok, BFp12* f) {
  switch (cfg / 2) {
#define X(ID, B, M) case ID: k_bls_check_miller<B, M><<<(unsigned)((n + B - 1) / B), B, 0, ctx->stream>>>(n, a1, a2, b1, b2, f, ok, b2_broadcast, pre_ok, ctx->d_flags); break;
    B2K_SPLIT_MILLER(X)
#undef X
    default: k_bls_check_miller<64, 8><<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>(n, a1, a2, b1, b2, f, ok, b2_broadcast, pre_ok, ctx->d_flags);
  }
  switch (cfg % 2) {
#define X(ID, B, M) case ID: k_bls_check_final<B, M><<<(unsigned)((n + B - 1) / B), B, 0, ctx->stream>>>(n, f, ok); break;
    B2K_SPLIT_FINAL(X)
#undef X
  }
}
inline void launch_pair_split(const b2k_ctx* ctx, int cfg, size_t n, const uint8_t* g1, const uint8_t* g2, uint8_t* gt, BFp12* f) {
  switch (cfg / 2) {
#define X(ID, B, M) case ID: k_bls_pair_miller<B, M><<<(unsigned)((n + B - 1) / B), B, 0, ctx->stream>>>(n, g1, g2, f, ctx->d_flags); break;
    B2K_SPLIT_MILLER(X)
#undef X
    default: k_bls_pair_miller<64, 8><<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>(n, g1, g2, f, ctx->d_flags);
  }
  switch (cfg % 2) {
#define X(ID, B, M) case ID: k_bls_pair_final<B, M><<<(unsigned)((n + B - 1) / B), B, 0, ctx->stream>>>(n, f, gt); break;
    B2K_SPLIT_FINAL(X)
#undef X
  }
}
#endif  // B2K_COMPACT_FIELD

// launch-bound variants: (threads per block, min blocks per SM) -> register cap 65536 / (threads * blocks)
#define B2K_PAIR_VARIANTS(X) X(0, 64, 4) X(1, 64, 8) X(2, 64, 6)   /* shapes (threads, min blocks per SM); (64, 4) = 255 registers wins every sweep (profiles/r01*, r02l) */
inline void launch_pair_v(const b2k_ctx* ctx, int variant, size_t n, const uint8_t* g1, const uint8_t* g2, uint8_t* gt) {
  switch (variant) {
#define X(ID, B, M) case ID: k_bls_pair<B, M><<<(unsigned)((n + B - 1) / B), B, 0, ctx->stream>>>(n, g1, g2, gt, ctx->d_flags); break;
    B2K_PAIR_VARIANTS(X)
#undef X
    default: k_bls_pair<64, 4><<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>(n, g1, g2, gt, ctx->d_flags);
  }
}
inline void launch_pairing_check_v(const b2k_ctx* ctx, int variant, size_t n, const uint8_t* a1, const uint8_t* a2, const uint8_t* b1,
                                   const uint8_t* b2, uint8_t* ok, int b2_broadcast, const uint8_t* pre_ok) {
  switch (variant) {
#define X(ID, B, M) case ID: k_bls_pairing_check<B, M><<<(unsigned)((n + B - 1) / B), B, 0, ctx->stream>>>(n, a1, a2, b1, b2, ok, b2_broadcast, pre_ok, ctx->d_flags); break;
    B2K_PAIR_VARIANTS(X)
#undef X
    default: k_bls_pairing_check<64, 4><<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>(n, a1, a2, b1, b2, ok, b2_broadcast, pre_ok, ctx->d_flags);
  }
}

}  // namespace b2k

// The pairing kernels exist in two code layouts, variant = shape + 4 * layout (b2k_set_pairing_variant):
//   layout 0 (variants 0..2, b2k_pairing.cu, the default): B2K_COMPACT_FIELD -- Fp and Fp2 products are out-of-line by-value calls
//   layout 1 (variants 4..6, b2k_pairing_inlined.cu): every field product inlined at its point of use
// (a third layout -- Fp product by value, Fp2 product out of line BY REFERENCE -- and a (64, 7) shape were measured and dropped:
//  81.4 ms against 75.5 ms per 65 536 checks, and (64, 7) compiles to the 128 registers of (64, 8); profiles/r02l_pairing_variants.txt)
// Every caller (b2k_pairing.cu, the bls.Verify paths of b2k_h2c.cu) launches through these two functions.
extern "C" void b2k_internal_launch_pair(const b2k_ctx* ctx, size_t n, const uint8_t* g1, const uint8_t* g2, uint8_t* gt);
extern "C" void b2k_internal_launch_pairing_check(const b2k_ctx* ctx, size_t n, const uint8_t* a1, const uint8_t* a2, const uint8_t* b1,
                                                  const uint8_t* b2, uint8_t* ok, int b2_broadcast, const uint8_t* pre_ok);
extern "C" void b2k_internal_launch_pair_inlined(const b2k_ctx* ctx, int variant, size_t n, const uint8_t* g1, const uint8_t* g2, uint8_t* gt);
extern "C" void b2k_internal_launch_pairing_check_inlined(const b2k_ctx* ctx, int variant, size_t n, const uint8_t* a1, const uint8_t* a2,
                                                          const uint8_t* b1, const uint8_t* b2, uint8_t* ok, int b2_broadcast,
                                                          const uint8_t* pre_ok);
