// coop_core.cuh -- the per-lane core of the warp-cooperative interpreter (coop_pairing.cuh): slot layout and one operation of one lane.
// No warp collectives and no kernels in here, so that the host emulation (tests/host_emul/emul_coop.cpp) compiles exactly this code.
#pragma once
#include "fp.cuh"
#include "fp_inv.cuh"

namespace b2k {
namespace coop {

// slots of one warp in shared memory, limb-major: word j of slot s at [j * NS + s] (NS odd: lanes on different slots hit different banks)
template <class P1, class P2> struct Layout {
  static constexpr int NS = (P1::SLOTS > P2::SLOTS ? P1::SLOTS : P2::SLOTS) | 1;
  static constexpr size_t BYTES = (size_t)P1::LIMBS * NS * 4;
};

template <int NS, class C>
B2K_D void slot_load(Fp<C>& r, const uint32_t* S, uint32_t s) {
#pragma unroll
  for (int j = 0; j < C::N; j++) r.v[j] = S[j * NS + s];
}
template <int NS, class C>
B2K_D void slot_store(uint32_t* S, uint32_t s, const Fp<C>& a) {
#pragma unroll
  for (int j = 0; j < C::N; j++) S[j * NS + s] = a.v[j];
}

// one round: decode the lane's word, load operands, compute, store, __syncwarp().  Operations: 1 MUL, 3 ADD, 4 SUB, 7 MULC, 8 INV
// (the generator encodes a^2, 2 a, -a as a * a, a + a, ZERO - a).  A round holds products only or additions / subtractions only, and
// the two kinds take SEPARATE, warp-uniform code paths: the product path calls out-of-line functions, and sharing variables with it made
// the compiler park the operands of every addition round on the stack.
// (split so that the host emulation, tests/host_emul/emul_coop.cpp, can run the per-lane part lane by lane: step_lane has no warp collective)
template <int NS, class C>
B2K_D void step_lane(uint32_t* S, uint32_t w, const uint32_t* consts, bool additive_round) {
  const uint32_t op = w >> 28, d = (w >> 18) & 511u, a = (w >> 9) & 511u, b = w & 511u;
  if (additive_round) {                                      // an addition / subtraction round
    if (op == 3 || op == 4) {
      Fp<C> x, y, z;
      slot_load<NS>(x, S, a);
      slot_load<NS>(y, S, b);
      fp_addsub(z, x, y, op == 4);                           // one instruction stream for both (fp.cuh)
      slot_store<NS>(S, d, z);
    }
  } else if (op != 0) {                                      // a product round (or the lone inversion)
    Fp<C> x, y, z;
    slot_load<NS>(x, S, a);
    if (op == 8) {
      fp_inv_bingcd(z, x);
    } else {
      if (op == 7) {
#pragma unroll
        for (int j = 0; j < C::N; j++) y.v[j] = consts[b * C::N + j];
      } else slot_load<NS>(y, S, b);
      fp_mul(z, x, y);
    }
    slot_store<NS>(S, d, z);
  }
}
}  // namespace coop
}  // namespace b2k
