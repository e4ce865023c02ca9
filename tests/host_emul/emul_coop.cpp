// Host emulation of the warp-cooperative interpreter (test infrastructure only; never loaded by the product): the generated program
// tables and kyber_b200/csrc/coop_core.cuh (slot layout, one operation of one lane: fp_mul / fp_addsub / inversion on Montgomery limbs,
// the Montgomery constants table) run lane after lane on the CPU.  A schedule in which a lane read a slot written in the SAME round
// would give a different answer here than in the read-all-then-write model of tools/gen_coop_pairing.py --check.
#include <cstring>
#include <vector>
#define __device__
#define __forceinline__ inline
#include "../../kyber_b200/csrc/constants.cuh"
#include "../../kyber_b200/csrc/coop_core.cuh"
#include "../../kyber_b200/csrc/coop_program_bls.inc"
#include "../../kyber_b200/csrc/coop_program_bn254.inc"
#include "../../kyber_b200/csrc/coop_program_bls_gt.inc"
using namespace b2k;

// inputs: PG::INPUTS field elements (plain integers, little-endian 32-bit limbs); out: 12 field elements (plain integers)
template <class PG, class C>
static void run_program(const uint32_t* in, uint32_t* out) {
  constexpr int NS = PG::SLOTS | 1, N = C::N;
  std::vector<uint32_t> S((size_t)N * NS, 0xEEEEEEEEu);
  for (int k = 0; k < PG::INPUTS; k++) {
    Fp<C> t, m;
    for (int j = 0; j < N; j++) t.v[j] = in[k * N + j];
    fp_to_mont(m, t);
    coop::slot_store<NS>(S.data(), k, m);
  }
  for (int j = 0; j < N; j++) { S[(size_t)j * NS + PG::ONE] = C::r1(j); S[(size_t)j * NS + PG::ZERO] = 0; }
  const uint32_t* prog = PG::prog();
  for (int r = 0; r < PG::ROUNDS; r++) {
    bool additive = false;
    for (int l = 0; l < 32; l++) { const uint32_t op = prog[r * 32 + l] >> 28; additive = additive || op == 3 || op == 4; }
    for (int l = 0; l < 32; l++) coop::step_lane<NS, C>(S.data(), prog[r * 32 + l], PG::consts(), additive);
  }
  for (int k = 0; k < 12; k++) {
    Fp<C> m, t;
    coop::slot_load<NS>(m, S.data(), PG::out()[k]);
    fp_from_mont(t, m);
    for (int j = 0; j < N; j++) out[k * N + j] = t.v[j];
  }
}

extern "C" {
void emul_coop_bls_p1(const uint32_t* in, uint32_t* out) { run_program<coop::BLS_P1, Bls381Fp>(in, out); }
void emul_coop_bls_p2(const uint32_t* in, uint32_t* out) { run_program<coop::BLS_P2, Bls381Fp>(in, out); }
void emul_coop_bn254_p1(const uint32_t* in, uint32_t* out) { run_program<coop::BN254_P1, Bn254Fp>(in, out); }
void emul_coop_bls_sqr(const uint32_t* in, uint32_t* out) { run_program<coop::BLS_SQR, Bls381Fp>(in, out); }
void emul_coop_bls_mul(const uint32_t* in, uint32_t* out) { run_program<coop::BLS_MUL, Bls381Fp>(in, out); }
}
