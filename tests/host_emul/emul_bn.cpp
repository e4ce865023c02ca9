// Host emulation of the bn254 pairing device code (test infrastructure only).
#include "../../kyber_b200/csrc/bn_pairing.cuh"
using namespace b2k;
extern "C" {
void emul_bn254_pair(const uint8_t* g1, const uint8_t* g2, uint8_t* gt384) {
  Affine<NFp> P; Affine<NFp2> Q;
  bn254_g1_load(P, g1); bn254_g2_load(Q, g2);
  NFp12 f, e;
  bn254_miller_loop<1>(f, &P, &Q);
  bn254_final_exponentiation(e, f);
  if (aff_is_inf(P) || aff_is_inf(Q)) fp12_set_one(e);
  bn254_gt_store(gt384, e);
}
int emul_bn254_pairing_check(const uint8_t* a1, const uint8_t* a2, const uint8_t* b1, const uint8_t* b2) {
  Affine<NFp> P[2]; Affine<NFp2> Q[2];
  bn254_g1_load(P[0], a1); bn254_g2_load(Q[0], a2);
  bn254_g1_load(P[1], b1); bn254_g2_load(Q[1], b2);
  fp_neg(P[1].y, P[1].y);
  NFp12 f, e;
  bn254_miller_loop<2>(f, P, Q);
  bn254_final_exponentiation(e, f);
  return fp12_is_one(e) ? 1 : 0;
}
}

extern "C" void emul_bn256_pair(const uint8_t* g1, const uint8_t* g2, uint8_t* gt384) {
  using PC = Bn256Pair;
  Affine<PFp<PC>> P; Affine<PFp2<PC>> Q;
  bn_g1_load<PC>(P, g1); bn_g2_load<PC>(Q, g2);
  PFp12<PC> f, e;
  bn_miller_loop<PC, 1>(f, &P, &Q);
  bn_final_exponentiation<PC>(e, f);
  if (aff_is_inf(P) || aff_is_inf(Q)) fp12_set_one(e);
  bn_gt_store<PC>(gt384, e);
}
