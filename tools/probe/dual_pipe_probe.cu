// dual_pipe_probe.cu -- the experiment that opens the next round: chains of mixed XYZZ additions (the inner loop of the MSM
// accumulate pass) with FP64_WARPS of every 4 warps on the FP64-form field (tools/probe/ec_dfma.cuh, FP64 pipe) and the others on
// the library's 12 x 32-bit field (IMAD.WIDE, FMA-heavy pipe).  Prints additions/s for the splits 0..4 and checks that every
// split gives the same buckets, limb for limb.  Operands come from a 256-point table (L1/L2 resident): compute-bound by design.
//   build: nvcc -O3 -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -o tools/probe/dual_pipe_probe tools/probe/dual_pipe_probe.cu
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../../kyber_b200/csrc/constants.cuh"
#include "../../kyber_b200/csrc/fp.cuh"
#include "../../kyber_b200/csrc/curves.cuh"
#include "msm_slice_fp64.cuh"

using namespace b2k;
using CV = Bls381G1;
using F = CV::F;
constexpr int TABLE = 256;

__global__ void k_table(Affine<F>* table) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= TABLE) return;
  Affine<F> g;
  CV::generator(g);
  Xyzz<F> acc;
  xyzz_set_inf(acc);
  for (int i = 0; i <= 3 * k + 1; i++) xyzz_madd(acc, acc, g, false);      // (3k + 2) G: distinct, never the doubling case below
  Affine<F> a;
  xyzz_to_affine(a, acc);
  table[k] = a;
}

template <int FP64_WARPS>
__global__ void __launch_bounds__(128, 3) k_chain(int iters, const Affine<F>* __restrict__ table, Xyzz<F>* __restrict__ out) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if ((int)(threadIdx.x >> 5) < FP64_WARPS) {
    dfma::Fp one;
    { F o; f_set_one(o); dfma::from_u32(one, o.v); }
    dfma::Xyzz acc;
    fp64_set_inf(acc);
    for (int i = 0; i < iters; i++) fp64_madd(acc, table[(tid * 7 + i * 13) & (TABLE - 1)], (i & 3) == 3, one);
    Xyzz<F> r;
    fp64_store_xyzz(r, acc);
    out[tid] = r;
  } else {
    Xyzz<F> acc;
    xyzz_set_inf(acc);
    for (int i = 0; i < iters; i++) xyzz_madd(acc, acc, table[(tid * 7 + i * 13) & (TABLE - 1)], (i & 3) == 3);
    out[tid] = acc;
  }
}

template <int FP64_WARPS>
static double run(int grid, int iters, const Affine<F>* table, Xyzz<F>* d_out, std::vector<Xyzz<F>>& h) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_chain<FP64_WARPS><<<grid, 128>>>(8, table, d_out);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_chain<FP64_WARPS><<<grid, 128>>>(iters, table, d_out);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  h.resize((size_t)grid * 128);
  cudaMemcpy(h.data(), d_out, h.size() * sizeof(Xyzz<F>), cudaMemcpyDeviceToHost);
  const double rate = (double)grid * 128 * iters / (ms * 1e-3);
  printf("FP64 warps per block of 4: %d   %.3f ms   %.3e additions/s   (%s)\n", FP64_WARPS, ms, rate, cudaGetErrorString(cudaGetLastError()));
  return rate;
}

int main() {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { printf("no device\n"); return 1; }
  const int grid = prop.multiProcessorCount * 3, iters = 512;
  Affine<F>* table; Xyzz<F>* d_out;
  cudaMalloc(&table, TABLE * sizeof(Affine<F>));
  cudaMalloc(&d_out, (size_t)grid * 128 * sizeof(Xyzz<F>));
  k_table<<<TABLE / 64, 64>>>(table);
  cudaDeviceSynchronize();
  printf("%s, %d SMs; %d blocks x 128 threads, %d mixed XYZZ additions per thread; IMAD-form ceiling 3.0e10 products/s = 3.0e9 additions/s\n",
         prop.name, prop.multiProcessorCount, grid, iters);
  std::vector<Xyzz<F>> ref, got;
  run<0>(grid, iters, table, d_out, ref);
  int bad = 0;
  run<1>(grid, iters, table, d_out, got); bad += memcmp(ref.data(), got.data(), ref.size() * sizeof(Xyzz<F>)) != 0;
  run<2>(grid, iters, table, d_out, got); bad += memcmp(ref.data(), got.data(), ref.size() * sizeof(Xyzz<F>)) != 0;
  run<3>(grid, iters, table, d_out, got); bad += memcmp(ref.data(), got.data(), ref.size() * sizeof(Xyzz<F>)) != 0;
  run<4>(grid, iters, table, d_out, got); bad += memcmp(ref.data(), got.data(), ref.size() * sizeof(Xyzz<F>)) != 0;
  printf(bad ? "MISMATCH between the forms in %d split(s)\n" : "all splits give identical buckets (%d mismatches)\n", bad);
  return bad != 0;
}
