// b2k_msm_inlined.cu -- the BLS12-381 G1 MSM pipeline and Point.Mul batch in the INLINED code layout, kept for A/B runs.
// The library is built with B2K_COMPACT_FIELD (fp.cuh): every field product outside the affine pair-tree rounds is a call to one
// out-of-line by-value body.  Measured on B200 (profiles/r02g_layout_ab.txt): MSM 2^20 pipelined 6.26-6.32 -> 5.95 ms
// (1.67e8 -> 1.76e8 muls/s), Point.Mul batch 78.2 -> 71.5 ms per 2^20 (1.34e7 -> 1.47e7 /s).  This unit compiles the same
// templates with the products inlined at every use (the round-1 layout); b2k_set_msm_layout(ctx, 1) selects it.
#undef B2K_COMPACT_FIELD
// the same templates are instantiated in the compact layout elsewhere: keep this unit's instantiations (host stubs are weak
// symbols the linker merges by name) in their own namespaces
#define b2k b2k_inlined
#define b2k_host b2k_inlined_host
#include <cuda_runtime.h>
#include "../../include/b2kyber.h"
#include "msm_host.cuh"

using namespace b2k;
using namespace b2k_host;

extern "C" int b2k_internal_bls12381_g1_mul_batch_dev_inlined(b2k_ctx* c, size_t n, const void* s, const void* p, void* o, int affine_out) {
  return affine_out ? mul_batch_dev<Bls381G1, true>(c, n, s, p, o) : mul_batch_dev<Bls381G1, false>(c, n, s, p, o);
}
extern "C" int b2k_internal_bls12381_g1_msm_dev_inlined(b2k_ctx* c, size_t n, const void* s, const void* p, void* o, int affine_out) {
  return msm_dev<Bls381G1>(c, n, s, p, o, affine_out);
}
