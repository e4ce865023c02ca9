// coop_pairing.cuh -- WARP-COOPERATIVE pairings (BLS12-381, bn254, bn256): one warp per pairing (or per 2-pair check), the 32 lanes
// execute the independent Fp operations of the pairing in lock step.
//
// Replaces, for SMALL batches, the same reference symbols as the one-per-thread kernels (kilic.Suite.Pair / ValidatePairing,
// pairing/bls12381/kilic/suite.go:57-75; pairing/bn254/suite.go:133-144; pairing/bn256/suite.go:99-109): those run one pairing per
// THREAD, so a single call of the one-at-a-time interface method costs one thread's latency (20-26 ms for ~20 000 Fp products in
// program order).  The dependency DEPTH of a pairing is only ~800 products: tools/gen_coop_pairing.py compiles the whole computation
// (Miller loop(s) + final exponentiation; the loop bits are public) into rounds of <= 32 independent Fp operations with their operands
// in slots of shared memory; this file is the interpreter.  A round: lane l decodes word [round][l] = (op, dst, a, b), loads its
// operands, computes, stores, __syncwarp().  Products and additive operations never share a round, so a round costs one product or
// one addition.  The program tables are validated on the host (formulas, schedule, slot allocation, encoding:
// tests/test_coop_program.py) and the kernels on the GPU against the batch kernels and the test reference (tests/test_gpu_coop_pairing.py).
//
// A translation unit defines B2K_COOP_BLS / B2K_COOP_BN254 / B2K_COOP_BN256 in front of this header to pull in the tables it uses
// (and includes bn_pairing.cuh itself for the two Barreto-Naehrig curves).
#pragma once
#include "pairing.cuh"
#include "curves.cuh"
#include "codec.cuh"
#include "kernels.cuh"
#include "fp_inv.cuh"
#include "coop_core.cuh"
#ifdef B2K_COOP_BLS
#include "coop_program_bls.inc"
#endif
#ifdef B2K_COOP_BN254
#include "coop_program_bn254.inc"
#endif
#ifdef B2K_COOP_BN256
#include "coop_program_bn256.inc"
#endif
#ifdef B2K_COOP_GT                       // Fp12 square / product programs of the GT group (b2k_gt.cu)
#include "coop_program_bls_gt.inc"
#include "coop_program_bn254_gt.inc"
#endif

namespace b2k {
namespace coop {

template <int NS, class C>
B2K_D void step(uint32_t* S, uint32_t w, const uint32_t* consts) {
  const uint32_t op = w >> 28;
  step_lane<NS, C>(S, w, consts, __any_sync(0xffffffffu, op == 3 || op == 4));
  __syncwarp();
}

// the interpreter proper: every lane of the warp calls it with the same program; the words of the next two rounds travel while this one
// runs (an addition round is shorter than an L2 round trip; fetching a group of 8 rounds ahead was measured slower: the unrolled body costs
// more than the latency it hides)
template <class PG, int NS, class C>
B2K_D void run(uint32_t* S, int lane) {
  const uint32_t* __restrict__ prog = PG::prog();
  const uint32_t* consts = PG::consts();
  constexpr int rounds = PG::ROUNDS;
  uint32_t w = prog[lane], w1 = rounds > 1 ? prog[32 + lane] : 0u;
  for (int r = 0; r < rounds; r++) {
    const uint32_t w2 = (r + 2 < rounds) ? prog[(size_t)(r + 2) * 32 + lane] : 0u;
    step<NS, C>(S, w, consts);
    w = w1; w1 = w2;
  }
}

// a curve for the cooperative kernels: CV1 / CV2 = the G1 / G2 operand codecs (curves.cuh, codec.cuh, bn256.cuh), F12 the target field,
// P1 / P2 the 1-pair and 2-pair programs
template <class CV1_, class CV2_, class C_, class F12_, class P1_, class P2_>
struct Curve {
  using CV1 = CV1_; using CV2 = CV2_; using C = C_; using F12 = F12_; using P1 = P1_; using P2 = P2_;
  using L = Layout<P1, P2>;
};

// operands of one element -> input slots; returns false for a malformed operand.  live[i] = neither member of pair i is infinity.
// Lane 2 i loads P_i, lane 2 i + 1 loads Q_i (each a full decode + range + curve check, as in the batch kernels).
template <class CU, int NPAIRS>
B2K_D bool load_inputs(uint32_t* S, int lane, const uint8_t* const* g1, const uint8_t* const* g2, bool negate_second, bool* live) {
  constexpr int NS = CU::L::NS;
  bool good = true, inf = false;
  if (lane < 2 * NPAIRS) {
    const int i = lane >> 1;
    if ((lane & 1) == 0) {
      Affine<typename CU::CV1::F> Pt;
      good = load_checked<typename CU::CV1>(Pt, g1[i]);
      inf = aff_is_inf(Pt);
      if (negate_second && i == 1) fp_neg(Pt.y, Pt.y);
      slot_store<NS>(S, 6 * i, Pt.x); slot_store<NS>(S, 6 * i + 1, Pt.y);
    } else {
      Affine<typename CU::CV2::F> Q;
      good = load_checked<typename CU::CV2>(Q, g2[i]);
      inf = aff_is_inf(Q);
      slot_store<NS>(S, 6 * i + 2, Q.x.c0); slot_store<NS>(S, 6 * i + 3, Q.x.c1);
      slot_store<NS>(S, 6 * i + 4, Q.y.c0); slot_store<NS>(S, 6 * i + 5, Q.y.c1);
    }
  }
  const unsigned bad = __ballot_sync(0xffffffffu, !good), infm = __ballot_sync(0xffffffffu, inf);
  for (int i = 0; i < NPAIRS; i++) live[i] = ((infm >> (2 * i)) & 3u) == 0;
  return bad == 0;
}

template <class CU>
B2K_D void write_consts(uint32_t* S, int lane, uint32_t one_slot, uint32_t zero_slot) {   // Montgomery one and zero into the program's pinned slots
  if (lane < CU::C::N) { S[lane * CU::L::NS + one_slot] = CU::C::r1(lane); S[lane * CU::L::NS + zero_slot] = 0u; }
}
template <class CU>
B2K_D void gather12(typename CU::F12& e, const uint32_t* S, const uint16_t* out) {   // flat order = the Fp12 memory layout (c0.c0.c0, c0.c0.c1, c0.c1.c0, ...)
  Fp<typename CU::C>* c = reinterpret_cast<Fp<typename CU::C>*>(&e);
  for (int k = 0; k < 12; k++) slot_load<CU::L::NS>(c[k], S, out[k]);
}

// e = the pairing of one element (one warp); valid in lane 0.  A malformed operand counts as infinity (e = 1, FLAG_POINT).
template <class CU>
B2K_D void pair_one(uint32_t* S, int lane, const uint8_t* g1, const uint8_t* g2, typename CU::F12& e, uint32_t* flags) {
  const uint8_t* p1[1] = {g1};
  const uint8_t* p2[1] = {g2};
  bool live[1];
  const bool good = load_inputs<CU, 1>(S, lane, p1, p2, false, live);
  if (!good && lane == 0) atomicOr(flags, FLAG_POINT);
  if (good && live[0]) {
    write_consts<CU>(S, lane, CU::P1::ONE, CU::P1::ZERO);
    __syncwarp();
    run<typename CU::P1, CU::L::NS, typename CU::C>(S, lane);
    if (lane == 0) gather12<CU>(e, S, CU::P1::out());
  } else if (lane == 0) fp12_set_one(e);
}

// ( e(a1, a2) == e(b1, b2) ) of one element; the boolean is valid in lane 0 (false for a malformed operand, FLAG_POINT raised)
template <class CU>
B2K_D bool check_one(uint32_t* S, int lane, const uint8_t* a1, const uint8_t* a2, const uint8_t* b1, const uint8_t* b2, uint32_t* flags) {
  constexpr int NS = CU::L::NS;
  const uint8_t* p1[2] = {a1, b1};
  const uint8_t* p2[2] = {a2, b2};
  bool live[2];
  const bool good = load_inputs<CU, 2>(S, lane, p1, p2, true, live);
  if (!good) { if (lane == 0) atomicOr(flags, FLAG_POINT); return false; }
  bool one = true;
  if (live[0] && live[1]) {
    write_consts<CU>(S, lane, CU::P2::ONE, CU::P2::ZERO);
    __syncwarp();
    run<typename CU::P2, NS, typename CU::C>(S, lane);
    if (lane == 0) { typename CU::F12 e; gather12<CU>(e, S, CU::P2::out()); one = fp12_is_one(e); }
  } else if (live[0] || live[1]) {                           // one pair has an infinity member and contributes 1: a 1-pair product
    __syncwarp();
    if (live[1] && lane < CU::C::N) {                        // move pair 1 into the input slots of the 1-pair program (limb-major: lane = limb)
      for (int s = 0; s < 6; s++) S[lane * NS + s] = S[lane * NS + 6 + s];
    }
    __syncwarp();
    write_consts<CU>(S, lane, CU::P1::ONE, CU::P1::ZERO);
    __syncwarp();
    run<typename CU::P1, NS, typename CU::C>(S, lane);
    if (lane == 0) { typename CU::F12 e; gather12<CU>(e, S, CU::P1::out()); one = fp12_is_one(e); }
  }
  return one;
}

#ifdef B2K_COOP_BLS
using Bls = Curve<Bls381G1, Bls381G2, Bls381Fp, BFp12, BLS_P1, BLS_P2>;
constexpr size_t SMEM_BYTES = Bls::L::BYTES;

// gt[i] = e(g1[i], g2[i])
static __global__ void __launch_bounds__(32, 16) k_coop_pair(size_t n, const uint8_t* __restrict__ g1, const uint8_t* __restrict__ g2,
                                                             uint8_t* __restrict__ gt, uint32_t* flags) {
  extern __shared__ __align__(16) uint32_t coop_sm[];
  const size_t i = blockIdx.x;
  if (i >= n) return;
  BFp12 e;
  pair_one<Bls>(coop_sm, threadIdx.x, g1 + 96 * i, g2 + 192 * i, e, flags);
  if (threadIdx.x == 0) gt_store(gt + 576 * i, e);
}

// ok[i] = ( e(a1[i], a2[i]) == e(b1[i], b2[i]) ); same arguments as k_bls_pairing_check
static __global__ void __launch_bounds__(32, 16) k_coop_pairing_check(size_t n, const uint8_t* __restrict__ a1, const uint8_t* __restrict__ a2,
                                                                      const uint8_t* __restrict__ b1, const uint8_t* __restrict__ b2,
                                                                      uint8_t* __restrict__ ok, int b2_broadcast,
                                                                      const uint8_t* __restrict__ pre_ok, uint32_t* flags) {
  extern __shared__ __align__(16) uint32_t coop_sm[];
  const size_t i = blockIdx.x;
  if (i >= n) return;
  if (pre_ok && !pre_ok[i]) { if (threadIdx.x == 0) ok[i] = 0; return; }
  const bool one = check_one<Bls>(coop_sm, threadIdx.x, a1 + 96 * i, a2 + 192 * i, b1 + ((b2_broadcast & 2) ? 0 : 96 * i),
                                  b2 + ((b2_broadcast & 1) ? 0 : 192 * i), flags);
  if (threadIdx.x == 0) ok[i] = one ? 1 : 0;
}
#endif

// the Barreto-Naehrig curves: 64-byte G1 and 128-byte G2 operands, 384-byte GT (bn_pairing.cuh); PC = Bn254Pair / Bn256Pair
#if defined(B2K_COOP_BN254) || defined(B2K_COOP_BN256)
template <class CU, class PC>
static __global__ void __launch_bounds__(32, 16) k_coop_bn_pair(size_t n, const uint8_t* __restrict__ g1, const uint8_t* __restrict__ g2,
                                                                uint8_t* __restrict__ gt, uint32_t* flags) {
  extern __shared__ __align__(16) uint32_t coop_sm[];
  const size_t i = blockIdx.x;
  if (i >= n) return;
  typename CU::F12 e;
  pair_one<CU>(coop_sm, threadIdx.x, g1 + 64 * i, g2 + 128 * i, e, flags);
  if (threadIdx.x == 0) bn_gt_store<PC>(gt + 384 * i, e);
}
template <class CU>
static __global__ void __launch_bounds__(32, 16) k_coop_bn_pairing_check(size_t n, const uint8_t* __restrict__ a1, const uint8_t* __restrict__ a2,
                                                                         const uint8_t* __restrict__ b1, const uint8_t* __restrict__ b2,
                                                                         uint8_t* __restrict__ ok, uint32_t* flags) {
  extern __shared__ __align__(16) uint32_t coop_sm[];
  const size_t i = blockIdx.x;
  if (i >= n) return;
  const bool one = check_one<CU>(coop_sm, threadIdx.x, a1 + 64 * i, a2 + 128 * i, b1 + 64 * i, b2 + 128 * i, flags);
  if (threadIdx.x == 0) ok[i] = one ? 1 : 0;
}
#endif

}  // namespace coop
}  // namespace b2k
