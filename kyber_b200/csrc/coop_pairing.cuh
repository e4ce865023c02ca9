// coop_pairing.cuh -- WARP-COOPERATIVE BLS12-381 pairing: one warp per pairing (or per 2-pair check), the 32 lanes execute the
// independent Fp operations of the pairing in lock step.
//
// Replaces, for SMALL batches, the same reference symbols as pairing_kernels.cuh (kilic.Suite.Pair / ValidatePairing,
// pairing/bls12381/kilic/suite.go:57-75): those kernels run one pairing per THREAD, so a single call of the one-at-a-time interface
// method costs one thread's latency (~37 ms for ~20 000 dependent-by-program-order Fp products).  The dependency DEPTH of a pairing
// is only ~800 products: tools/gen_coop_pairing.py compiles the whole computation (Miller loop(s) + final exponentiation, loop bits of
// |x| are public) into rounds of <= 32 independent Fp operations with their operands in slots of shared memory; this file is the
// interpreter.  A round: lane l decodes word [round][l] = (op, dst, a, b), loads its operands, computes, stores, __syncwarp().
// Products (MUL / SQR / MULC) and additive operations never share a round, so a round costs one product or one addition.
// The program tables are validated on the host (formulas, schedule, slot allocation, encoding: tests/test_coop_program.py) and the
// kernels on the GPU against the batch kernels and the test reference (tests/test_gpu_coop_pairing.py).
#pragma once
#include "pairing.cuh"
#include "curves.cuh"
#include "codec.cuh"
#include "kernels.cuh"
#include "fp_inv.cuh"
#include "coop_program.inc"

namespace b2k {
namespace coop {

constexpr int NS = (P1_SLOTS > P2_SLOTS ? P1_SLOTS : P2_SLOTS) | 1;      // slots per warp (odd stride: lanes on different slots hit different banks)
constexpr size_t SMEM_BYTES = (size_t)12 * NS * 4;                       // limb-major: word j of slot s at [j * NS + s]

B2K_D void slot_load(BFp& r, const uint32_t* S, uint32_t s) {
#pragma unroll
  for (int j = 0; j < 12; j++) r.v[j] = S[j * NS + s];
}
B2K_D void slot_store(uint32_t* S, uint32_t s, const BFp& a) {
#pragma unroll
  for (int j = 0; j < 12; j++) S[j * NS + s] = a.v[j];
}

// one round: decode the lane's word, load operands, compute, store, __syncwarp().  Operations: 1 MUL, 3 ADD, 4 SUB, 7 MULC, 8 INV
// (the generator encodes a^2, 2 a, -a as a * a, a + a, ZERO - a).  A round holds products only or additions / subtractions only, and
// the two kinds take SEPARATE, warp-uniform code paths: the product path calls out-of-line functions, and sharing variables with it made
// the compiler park the operands of every addition round on the stack.
B2K_D void step(uint32_t* S, uint32_t w) {
  const uint32_t op = w >> 28, d = (w >> 18) & 511u, a = (w >> 9) & 511u, b = w & 511u;
  const bool additive = (op == 3 || op == 4);
  if (__any_sync(0xffffffffu, additive)) {                   // an addition / subtraction round
    if (additive) {
      BFp x, y, z;
      slot_load(x, S, a);
      slot_load(y, S, b);
      fp_addsub(z, x, y, op == 4);                           // one instruction stream for both (fp.cuh)
      slot_store(S, d, z);
    }
  } else if (op != 0) {                                      // a product round (or the lone inversion)
    BFp x, y, z;
    slot_load(x, S, a);
    if (op == 8) {
      fp_inv_bingcd(z, x);
    } else {
      if (op == 7) {
#pragma unroll
        for (int j = 0; j < 12; j++) y.v[j] = CONSTS[b][j];
      } else slot_load(y, S, b);
      fp_mul(z, x, y);
    }
    slot_store(S, d, z);
  }
  __syncwarp();
}

// the interpreter proper: every lane of the warp calls it with the same program; the words of the next two rounds travel while this one
// runs (an addition round is shorter than an L2 round trip; fetching a group of 8 rounds ahead was measured slower: the unrolled body costs
// more than the latency it hides)
B2K_D void run(uint32_t* S, const uint32_t* __restrict__ prog, int rounds, int lane) {
  uint32_t w = prog[lane], w1 = rounds > 1 ? prog[32 + lane] : 0u;
  for (int r = 0; r < rounds; r++) {
    const uint32_t w2 = (r + 2 < rounds) ? prog[(size_t)(r + 2) * 32 + lane] : 0u;
    step(S, w);
    w = w1; w1 = w2;
  }
}

// operands of one element -> input slots; returns false for a malformed operand.  live[i] = neither member of pair i is infinity.
// Lane 2 i loads P_i, lane 2 i + 1 loads Q_i (each a full decode + range + curve check, as in the batch kernels).
template <int NPAIRS>
B2K_D bool load_inputs(uint32_t* S, int lane, const uint8_t* const* g1, const uint8_t* const* g2, bool negate_second, bool* live) {
  bool good = true, inf = false;
  if (lane < 2 * NPAIRS) {
    const int i = lane >> 1;
    if ((lane & 1) == 0) {
      Affine<BFp> Pt;
      good = load_checked<Bls381G1>(Pt, g1[i]);
      inf = aff_is_inf(Pt);
      if (negate_second && i == 1) fp_neg(Pt.y, Pt.y);
      slot_store(S, 6 * i, Pt.x); slot_store(S, 6 * i + 1, Pt.y);
    } else {
      Affine<BFp2> Q;
      good = load_checked<Bls381G2>(Q, g2[i]);
      inf = aff_is_inf(Q);
      slot_store(S, 6 * i + 2, Q.x.c0); slot_store(S, 6 * i + 3, Q.x.c1);
      slot_store(S, 6 * i + 4, Q.y.c0); slot_store(S, 6 * i + 5, Q.y.c1);
    }
  }
  const unsigned bad = __ballot_sync(0xffffffffu, !good), infm = __ballot_sync(0xffffffffu, inf);
  for (int i = 0; i < NPAIRS; i++) live[i] = ((infm >> (2 * i)) & 3u) == 0;
  return bad == 0;
}

B2K_D void write_consts(uint32_t* S, int lane, uint32_t one_slot, uint32_t zero_slot) {   // Montgomery one and zero into the program's pinned slots
  if (lane < 12) { S[lane * NS + one_slot] = Bls381Fp::r1(lane); S[lane * NS + zero_slot] = 0u; }
}
B2K_D void gather12(BFp12& e, const uint32_t* S, const uint16_t* out) {     // flat order = BFp12's memory layout (c0.c0.c0, c0.c0.c1, c0.c1.c0, ...)
  BFp* c = reinterpret_cast<BFp*>(&e);
  for (int k = 0; k < 12; k++) slot_load(c[k], S, out[k]);
}

// gt[i] = e(g1[i], g2[i])
static __global__ void __launch_bounds__(32, 16) k_coop_pair(size_t n, const uint8_t* __restrict__ g1, const uint8_t* __restrict__ g2,
                                                         uint8_t* __restrict__ gt, uint32_t* flags) {
  extern __shared__ __align__(16) uint32_t coop_sm[];
  const int lane = threadIdx.x;
  const size_t i = blockIdx.x;
  if (i >= n) return;
  const uint8_t* p1[1] = {g1 + 96 * i};
  const uint8_t* p2[1] = {g2 + 192 * i};
  bool live[1];
  const bool good = load_inputs<1>(coop_sm, lane, p1, p2, false, live);
  if (!good && lane == 0) atomicOr(flags, FLAG_POINT);      // a malformed operand counts as infinity: e = 1
  BFp12 e;
  if (good && live[0]) {
    write_consts(coop_sm, lane, P1_ONE, P1_ZERO);
    __syncwarp();
    run(coop_sm, P1_PROG, P1_ROUNDS, lane);
    if (lane == 0) gather12(e, coop_sm, P1_OUT);
  } else if (lane == 0) fp12_set_one(e);
  if (lane == 0) gt_store(gt + 576 * i, e);
}

// ok[i] = ( e(a1[i], a2[i]) == e(b1[i], b2[i]) ); same arguments as k_bls_pairing_check
static __global__ void __launch_bounds__(32, 16) k_coop_pairing_check(size_t n, const uint8_t* __restrict__ a1, const uint8_t* __restrict__ a2,
                                                                  const uint8_t* __restrict__ b1, const uint8_t* __restrict__ b2,
                                                                  uint8_t* __restrict__ ok, int b2_broadcast,
                                                                  const uint8_t* __restrict__ pre_ok, uint32_t* flags) {
  extern __shared__ __align__(16) uint32_t coop_sm[];
  const int lane = threadIdx.x;
  const size_t i = blockIdx.x;
  if (i >= n) return;
  if (pre_ok && !pre_ok[i]) { if (lane == 0) ok[i] = 0; return; }
  const uint8_t* p1[2] = {a1 + 96 * i, b1 + ((b2_broadcast & 2) ? 0 : 96 * i)};
  const uint8_t* p2[2] = {a2 + 192 * i, b2 + ((b2_broadcast & 1) ? 0 : 192 * i)};
  bool live[2];
  const bool good = load_inputs<2>(coop_sm, lane, p1, p2, true, live);
  if (!good) { if (lane == 0) { atomicOr(flags, FLAG_POINT); ok[i] = 0; } return; }   // malformed operand: the check fails
  bool one = true;
  if (live[0] && live[1]) {
    write_consts(coop_sm, lane, P2_ONE, P2_ZERO);
    __syncwarp();
    run(coop_sm, P2_PROG, P2_ROUNDS, lane);
    if (lane == 0) { BFp12 e; gather12(e, coop_sm, P2_OUT); one = fp12_is_one(e); }
  } else if (live[0] || live[1]) {                           // one pair has an infinity member and contributes 1: a 1-pair product
    __syncwarp();
    if (live[1] && lane < 12) {                              // move pair 1 into the input slots of the 1-pair program (limb-major: lane = limb)
      for (int s = 0; s < 6; s++) coop_sm[lane * NS + s] = coop_sm[lane * NS + 6 + s];
    }
    __syncwarp();
    write_consts(coop_sm, lane, P1_ONE, P1_ZERO);
    __syncwarp();
    run(coop_sm, P1_PROG, P1_ROUNDS, lane);
    if (lane == 0) { BFp12 e; gather12(e, coop_sm, P1_OUT); one = fp12_is_one(e); }
  }
  if (lane == 0) ok[i] = one ? 1 : 0;
}

}  // namespace coop
}  // namespace b2k
