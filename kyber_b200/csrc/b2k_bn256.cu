// b2k_bn256.cu -- C ABI entry points for bn256 G1/G2 Point.Mul batches and MSM.
#include "msm_host.cuh"
#include "bn256.cuh"
using namespace b2k_host;
extern "C" {
int b2k_bn256_g1_mul_batch(b2k_ctx* c, size_t n, const uint8_t* s, const uint8_t* p, uint8_t* o) { return mul_batch_host<Bn256G1, false>(c, n, s, p, o); }
int b2k_bn256_g1_msm(b2k_ctx* c, size_t n, const uint8_t* s, const uint8_t* p, uint8_t* o) { return msm_host<Bn256G1>(c, n, s, p, o); }
int b2k_bn256_g2_mul_batch(b2k_ctx* c, size_t n, const uint8_t* s, const uint8_t* p, uint8_t* o) { return mul_batch_host<Bn256G2, false>(c, n, s, p, o); }
int b2k_bn256_g2_msm(b2k_ctx* c, size_t n, const uint8_t* s, const uint8_t* p, uint8_t* o) { return msm_host<Bn256G2>(c, n, s, p, o); }
}  // extern "C"
