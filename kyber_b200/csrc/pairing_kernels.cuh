// pairing_kernels.cuh -- __global__ wrappers for batched pairings (one pairing / check per thread).
#pragma once
#include "pairing.cuh"
#include "curves.cuh"

namespace b2k {

// gt[i] = e(g1[i], g2[i])           replaces n x Suite.Pair (kilic/suite.go:70-75)
static __global__ void __launch_bounds__(64) k_bls_pair(size_t n, const uint8_t* __restrict__ g1, const uint8_t* __restrict__ g2,
                                                  uint8_t* __restrict__ gt) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<BFp> P;
  Affine<BFp2> Q;
  Bls381G1::load(P, g1 + 96 * i);
  g2_load(Q, g2 + 192 * i);
  BFp12 f, e;
  miller_loop<1>(f, &P, &Q);
  final_exponentiation(e, f);
  gt_store(gt + 576 * i, e);
}

// ok[i] = ( e(a1[i], a2[i]) == e(b1[i], b2[i]) )     replaces n x Suite.ValidatePairing
// (kilic/suite.go:57-68): one 2-pair Miller loop (second pair negated) + one final exponentiation.
static __global__ void __launch_bounds__(64) k_bls_pairing_check(size_t n, const uint8_t* __restrict__ a1,
                                                           const uint8_t* __restrict__ a2,
                                                           const uint8_t* __restrict__ b1,
                                                           const uint8_t* __restrict__ b2, uint8_t* __restrict__ ok,
                                                           int b2_broadcast, const uint8_t* __restrict__ pre_ok) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (pre_ok && !pre_ok[i]) { ok[i] = 0; return; }       // an operand failed UnmarshalBinary upstream
  Affine<BFp> P[2];
  Affine<BFp2> Q[2];
  Bls381G1::load(P[0], a1 + 96 * i);
  g2_load(Q[0], a2 + 192 * i);
  Bls381G1::load(P[1], b1 + 96 * i);
  g2_load(Q[1], b2 + (b2_broadcast ? 0 : 192 * i));
  fp_neg(P[1].y, P[1].y);
  BFp12 f, e;
  miller_loop<2>(f, P, Q);
  final_exponentiation(e, f);
  ok[i] = fp12_is_one(e) ? 1 : 0;
}

}  // namespace b2k
