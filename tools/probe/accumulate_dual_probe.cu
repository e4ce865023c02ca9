// Build-only prototype of next round's dual-pipe accumulate kernel: the balanced-slice pass of the MSM (kernels.cuh:
// k_msm_accumulate_slices) with the warps of a block split between the IMAD form (FMA-heavy pipe) and the FP64 form
// (tools/probe/msm_slice_fp64.cuh).  FP64_WARPS of every 4 warps take the FP64 body.  Both bodies are exact and interchangeable
// (tests/test_dfma_model.py runs them side by side under host emulation); what is left for the GPU is the split ratio and the
// register budget.  nvcc -gencode arch=compute_100a,code=sm_100a -I. accumulate_dual_probe.cu
#include <cstdio>
#include "../../kyber_b200/csrc/constants.cuh"
#include "../../kyber_b200/csrc/fp.cuh"
#include "msm_slice_fp64.cuh"

namespace b2k {
template <class CV, int FP64_WARPS>
__global__ void __launch_bounds__(128, 3) k_msm_accumulate_slices_dual(uint32_t nslices, uint32_t L, uint32_t total,
                                                                       const Affine<typename CV::F>* __restrict__ pts,
                                                                       const uint32_t* __restrict__ offs,
                                                                       const uint32_t* __restrict__ entries,
                                                                       Xyzz<typename CV::F>* __restrict__ buckets,
                                                                       Xyzz<typename CV::F>* __restrict__ spart) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nslices) return;
  if ((int)(threadIdx.x >> 5) < FP64_WARPS) msm_accumulate_slice_fp64<CV>(j, L, total, pts, offs, entries, buckets, spart);
  else msm_accumulate_slice<CV>(j, L, total, pts, offs, entries, buckets, spart);
}
template __global__ void k_msm_accumulate_slices_dual<Bls381G1, 2>(uint32_t, uint32_t, uint32_t, const Affine<Bls381G1::F>*, const uint32_t*,
                                                                   const uint32_t*, Xyzz<Bls381G1::F>*, Xyzz<Bls381G1::F>*);
}  // namespace b2k
int main() { printf("build-only\n"); return 0; }
