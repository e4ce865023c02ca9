// b2k_bn254.cu -- C ABI entry points for bn254 G1 (same kernel templates, 8-limb field).
#include "msm_host.cuh"
using namespace b2k_host;
extern "C" {
int b2k_bn254_g1_mul_batch(b2k_ctx* c, size_t n, const uint8_t* s, const uint8_t* p, uint8_t* o) { return mul_batch_host<Bn254G1, false>(c, n, s, p, o); }
int b2k_bn254_g1_msm(b2k_ctx* c, size_t n, const uint8_t* s, const uint8_t* p, uint8_t* o) { return msm_host<Bn254G1>(c, n, s, p, o); }
int b2k_bn254_g1_msm_dev(b2k_ctx* c, size_t n, const void* s, const void* p, void* o) { return msm_dev<Bn254G1>(c, n, s, p, o); }
}  // extern "C"
