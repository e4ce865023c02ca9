// Build-only: the backward kernel of the affine pair-tree rounds (kernels.cuh: k_pt_backward; 41 % of a pipelined MSM) instantiated on
// the FP64-form field.  Storage here is Affine<FpD> (8 doubles per coordinate); next round separates storage (12 x u32) from compute.
#include <cstdio>
#include "../../kyber_b200/csrc/constants.cuh"
#include "../../kyber_b200/csrc/fp.cuh"
#include "fpd_overloads.cuh"
#include "../../kyber_b200/csrc/msm_affine.cuh"
namespace b2k {
template <class CV, bool FIRST>
__global__ void __launch_bounds__(128, 3) k_pt_backward_fp64(uint32_t B, uint32_t total, const Affine<typename CV::F>* __restrict__ in,
                                                             const uint32_t* __restrict__ entries, const uint32_t* __restrict__ offs_in,
                                                             const uint32_t* __restrict__ offs_out, const typename CV::F* __restrict__ pre,
                                                             const typename CV::F* __restrict__ accs, Affine<typename CV::F>* __restrict__ out) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  msm_pairtree_backward<CV, FIRST>(t, B, gridDim.x * blockDim.x, total, in, entries, offs_in, offs_out, pre, accs, out);
}
template __global__ void k_pt_backward_fp64<Bls381G1D, false>(uint32_t, uint32_t, const Affine<FpD>*, const uint32_t*, const uint32_t*,
                                                              const uint32_t*, const FpD*, const FpD*, Affine<FpD>*);
}  // namespace b2k
int main() { printf("build-only\n"); return 0; }
