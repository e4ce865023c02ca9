// pipe_probe.cu -- raw issue rates of the integer-multiply (IMAD.WIDE) and FP64 (DFMA) pipes of one SM, alone and together.
// Question: can a 381-bit field multiplication be split between the two pipes (or warps be specialised) to beat the
// IMAD-only bound of the MSM bucket-accumulation kernel?   build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int CH = 8;

__device__ __forceinline__ uint64_t madwide(uint32_t a, uint32_t b, uint64_t c) {
  uint64_t r;
  asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
  return r;
}
__device__ __forceinline__ double dfma(double a, double b, double c) {
  double r;
  asm volatile("fma.rz.f64 %0, %1, %2, %3;" : "=d"(r) : "d"(a), "d"(b), "d"(c));
  return r;
}

// mode 0: imad only, 1: dfma only, 2: both interleaved in every thread, 3: even warps imad / odd warps dfma
__global__ void __launch_bounds__(256) k_probe(int mode, uint64_t* out, uint32_t seed) {
  uint64_t acc[CH];
  double dacc[CH];
  const uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
  const double da = 1.0 + 1e-9 * threadIdx.x, db = 1.0 - 1e-9 * blockIdx.x;
#pragma unroll
  for (int j = 0; j < CH; j++) { acc[j] = j; dacc[j] = j; }
  const bool do_i = mode == 0 || mode == 2 || (mode == 3 && ((threadIdx.x >> 5) & 1) == 0);
  const bool do_d = mode == 1 || mode == 2 || (mode == 3 && ((threadIdx.x >> 5) & 1) == 1);
  if (do_i && do_d) {
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
      for (int j = 0; j < CH; j++) { acc[j] = madwide(a, b, acc[j]); dacc[j] = dfma(da, dacc[j], db); }
    }
  } else if (do_i) {
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
      for (int j = 0; j < CH; j++) acc[j] = madwide(a, b, acc[j]);
#pragma unroll
      for (int j = 0; j < CH; j++) acc[j] = madwide(a, b, acc[j]);
    }
  } else {
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
      for (int j = 0; j < CH; j++) dacc[j] = dfma(da, dacc[j], db);
#pragma unroll
      for (int j = 0; j < CH; j++) dacc[j] = dfma(da, dacc[j], db);
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int j = 0; j < CH; j++) s += acc[j] + (uint64_t)__double_as_longlong(dacc[j]);
  if (s == 0x1234567) out[0] = s;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  uint64_t* out;
  cudaMalloc(&out, 8);
  const int blocks = p.multiProcessorCount * 4;
  const char* names[4] = {"imad.wide only", "dfma only", "both, every thread", "even warps imad / odd warps dfma"};
  for (int mode = 0; mode < 4; mode++) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_probe<<<blocks, 256>>>(mode, out, 7);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) k_probe<<<blocks, 256>>>(mode, out, 7 + r);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 5;
    // thread-level ops per kernel
    const double threads = (double)blocks * 256;
    double iops = 0, dops = 0;
    if (mode == 0) iops = threads * ITERS * 2.0 * CH;
    if (mode == 1) dops = threads * ITERS * 2.0 * CH;
    if (mode == 2) { iops = threads * ITERS * 1.0 * CH; dops = iops; }
    if (mode == 3) { iops = threads / 2 * ITERS * 2.0 * CH; dops = iops; }
    const double cyc = ms * 1e-3 * clk * 1e3;
    printf("%-36s %8.3f ms  imad/clk/SM %6.1f  dfma/clk/SM %6.1f  (clock attr %d kHz)\n", names[mode], ms,
           iops / cyc / p.multiProcessorCount, dops / cyc / p.multiProcessorCount, clk);
  }
  return 0;
}
