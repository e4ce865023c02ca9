// b2k_msm_compact.cu -- the BLS12-381 G1 MSM pipeline in the COMPACT code layout (fp.cuh: B2K_COMPACT_FIELD), for A/B runs
// against the inlined layout of b2k_api.cu: every field product of the bucket kernels is a call to one out-of-line by-value
// body (the XYZZ slice kernel has 53 KB of straight-line code per loop iteration otherwise and spends 22 % of its issue
// cycles waiting for instructions, profiles/r01i_accumulate_ncu_details.txt).  Selected with b2k_set_msm_layout(ctx, 1).
#define B2K_COMPACT_FIELD 1
#define b2k b2k_compact
#define b2k_host b2k_compact_host
#include <cuda_runtime.h>
#include "../../include/b2kyber.h"
#include "msm_host.cuh"

using namespace b2k;
using namespace b2k_host;

extern "C" int b2k_internal_bls12381_g1_msm_dev_compact(b2k_ctx* c, size_t n, const void* s, const void* p, void* o, int affine_out) {
  return msm_dev<Bls381G1>(c, n, s, p, o, affine_out);
}
