// b2k_pairing_compact.cu -- the BLS12-381 pairing kernels in the COMPACT code layout (fp.cuh: B2K_COMPACT_FIELD): every field
// product is a call to one out-of-line by-value body instead of ~330 inlined instructions, so that the hot loop fits the
// instruction caches (ncu, profiles/r02d_pairing_check_ncu_details.txt: the inlined layout stalls 35 % of its issue cycles
// on instruction fetch with 530 KB of code).  Same source (pairing.cuh), same results; selected with
// b2k_set_pairing_variant(ctx, 3..5) = the launch shapes 0..2 of the inlined layout.
#define B2K_COMPACT_FIELD 1
// the same templates are instantiated with the inlined layout in other translation units: keep this unit's instantiations
// (host stubs are weak symbols the linker would merge by name) in their own namespaces
#define b2k b2k_compact
#define b2k_host b2k_compact_host
#include <cuda_runtime.h>
#include "../../include/b2kyber.h"
#include "b2k_ctx.h"
#include "pairing_kernels.cuh"

using namespace b2k;

extern "C" {
void b2k_internal_launch_pair_compact(const b2k_ctx* ctx, int variant, size_t n, const uint8_t* g1, const uint8_t* g2, uint8_t* gt) {
  launch_pair_v(ctx, variant, n, g1, g2, gt);
}
void b2k_internal_launch_pairing_check_compact(const b2k_ctx* ctx, int variant, size_t n, const uint8_t* a1, const uint8_t* a2, const uint8_t* b1,
                                               const uint8_t* b2, uint8_t* ok, int b2_broadcast, const uint8_t* pre_ok) {
  launch_pairing_check_v(ctx, variant, n, a1, a2, b1, b2, ok, b2_broadcast, pre_ok);
}
}  // extern "C"
