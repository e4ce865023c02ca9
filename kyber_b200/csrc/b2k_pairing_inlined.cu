// b2k_pairing_inlined.cu -- the BLS12-381 pairing kernels in the INLINED code layout (every field product ~330 instructions at
// its point of use: 530 KB of code per kernel), kept for A/B runs.  ncu (profiles/r02d_pairing_check_ncu_details.txt) showed that
// layout stalled 35 % of its issue cycles on instruction fetch; the library default is the compact layout (B2K_COMPACT_FIELD,
// fp.cuh), measured 113.4 -> 83.5 ms per 65 536 checks (1.16e6 -> 1.57e6 pairings/s, profiles/r02g_layout_ab.txt).
// b2k_set_pairing_variant(ctx, 4..6) selects the launch shapes 0..2 in this layout.
#undef B2K_COMPACT_FIELD
#define b2k b2k_inlined
#define b2k_host b2k_inlined_host
#include <cuda_runtime.h>
#include "../../include/b2kyber.h"
#include "b2k_ctx.h"
#include "pairing_kernels.cuh"

using namespace b2k;

extern "C" {
void b2k_internal_launch_pair_inlined(const b2k_ctx* ctx, int variant, size_t n, const uint8_t* g1, const uint8_t* g2, uint8_t* gt) {
  launch_pair_v(ctx, variant, n, g1, g2, gt);
}
void b2k_internal_launch_pairing_check_inlined(const b2k_ctx* ctx, int variant, size_t n, const uint8_t* a1, const uint8_t* a2, const uint8_t* b1,
                                               const uint8_t* b2, uint8_t* ok, int b2_broadcast, const uint8_t* pre_ok) {
  launch_pairing_check_v(ctx, variant, n, a1, a2, b1, b2, ok, b2_broadcast, pre_ok);
}
}  // extern "C"
