// b2k_g1_mul.cu -- BLS12-381 G1 batched Point.Mul entry points (their own translation unit: the k_mul_batch variants
// compile in parallel with the MSM pipeline of b2k_api.cu).
#include <cuda_runtime.h>
#include "../../include/b2kyber.h"
#include "msm_host.cuh"

using namespace b2k;
using namespace b2k_host;

extern "C" {

int b2k_bls12381_g1_mul_batch(b2k_ctx* c, size_t n, const uint8_t* s, const uint8_t* p, uint8_t* o) { return mul_batch_host<Bls381G1, false>(c, n, s, p, o); }
int b2k_bls12381_g1_mul_batch_affine(b2k_ctx* c, size_t n, const uint8_t* s, const uint8_t* p, uint8_t* o) { return mul_batch_host<Bls381G1, true>(c, n, s, p, o); }
int b2k_bls12381_g1_mul_batch_dev(b2k_ctx* c, size_t n, const void* s, const void* p, void* o) {
  if (c && c->msm_layout == 1) return b2k_internal_bls12381_g1_mul_batch_dev_inlined(c, n, s, p, o, 0);      // A/B: inlined code layout
  return mul_batch_dev<Bls381G1, false>(c, n, s, p, o);
}
int b2k_bls12381_g1_mul_batch_affine_dev(b2k_ctx* c, size_t n, const void* s, const void* p, void* o) {
  if (c && c->msm_layout == 1) return b2k_internal_bls12381_g1_mul_batch_dev_inlined(c, n, s, p, o, 1);
  return mul_batch_dev<Bls381G1, true>(c, n, s, p, o);
}

}  // extern "C"
