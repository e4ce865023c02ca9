// imad_probe.cu -- which IMAD.WIDE forms run at 2 and which at 4 cycles per warp instruction on B200?
//   A: 196 products a[i]*b[j] (all operands distinct registers) into 28 64-bit column accumulators, no carries
//   B: the same with b[j] compile-time constants (immediates)
//   C: 14 products per step into ONE accumulator pair chain each (acc_k += a_i * b_i), 14 independent chains
//   D: saturated-style row: a_i * b_j with carry chain (mad.lo.cc / madc.hi.cc pairs -> IMAD.WIDE.U32.X)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(128) k(int iters, uint32_t* out, uint32_t seed) {
  uint32_t a[14], b[14];
  uint64_t t[28];
#pragma unroll
  for (int j = 0; j < 14; j++) { a[j] = (seed * 3 + j + threadIdx.x) & 0x1fffffffu; b[j] = (seed + 7 * j + blockIdx.x) & 0x1fffffffu; }
#pragma unroll
  for (int k2 = 0; k2 < 28; k2++) t[k2] = k2;
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 14; i++)
#pragma unroll
        for (int j = 0; j < 14; j++) t[i + j] += (uint64_t)a[i] * b[j];
    } else if (MODE == 1) {
      constexpr uint32_t cb[14] = {0x1fffaaabu, 0x0ff7ffffu, 0x14ffffeeu, 0x17fffd62u, 0x0f6241eau, 0x09507b58u, 0x0afd9cc3u,
                                   0x109e70a2u, 0x1764774bu, 0x121a5d66u, 0x12c6e9edu, 0x12ffcd34u, 0x00111ea3u, 0x0000000du};
#pragma unroll
      for (int i = 0; i < 14; i++)
#pragma unroll
        for (int j = 0; j < 14; j++) t[i + j] += (uint64_t)a[i] * cb[j];
    } else if (MODE == 2) {
#pragma unroll
      for (int r = 0; r < 14; r++)
#pragma unroll
        for (int i = 0; i < 14; i++) t[i] += (uint64_t)a[i] * b[(i + r) % 14];
    } else {
      // carry-chained rows: (lo,hi) of a[i]*b[j] added into 32-bit limb pairs with carries
      uint32_t* w = reinterpret_cast<uint32_t*>(t);     // 56 words
#pragma unroll
      for (int i = 0; i < 14; i++) {
        uint32_t lo, hi;
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(w[i]), "+r"(w[i + 1]) : "r"(a[i]), "r"(b[0]));
#pragma unroll
        for (int j = 2; j < 14; j += 2)
          asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(w[i + j]), "+r"(w[i + j + 1]) : "r"(a[i]), "r"(b[j]));
        asm volatile("addc.u32 %0, %0, 0;" : "+r"(w[i + 14]));
        (void)lo; (void)hi;
      }
    }
    // keep the operands moving so that nothing is loop-invariant
#pragma unroll
    for (int j = 0; j < 14; j++) { a[j] = (a[j] + (uint32_t)t[j]) & 0x1fffffffu; }
  }
  uint64_t s = 0;
#pragma unroll
  for (int k2 = 0; k2 < 28; k2++) s ^= t[k2];
  if (s == 0x12345) out[0] = (uint32_t)s;
}

template <int MODE>
static void run(const char* name, int nimad, int sms, int clk_khz, uint32_t* out) {
  const int iters = 4096, bps = 4;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms * bps, 128>>>(64, out, 5);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<sms * bps, 128>>>(iters, out, 9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double warps_per_smsp = bps * 4 / 4.0;
  const double cyc = ms * 1e-3 * clk_khz * 1e3;
  const double cyc_per_warp_imad = cyc / (iters * (double)nimad * warps_per_smsp);
  printf("%-44s %8.3f ms   %5.2f cycles per warp IMAD (per SM sub-partition)\n", name, ms, cyc_per_warp_imad);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  uint32_t* out;
  cudaMalloc(&out, 8);
  run<0>("A 196 x IMAD.WIDE, register operands", 196, p.multiProcessorCount, clk, out);
  run<1>("B 196 x IMAD.WIDE, immediate multiplicand", 196, p.multiProcessorCount, clk, out);
  run<2>("C 196 x IMAD.WIDE, 14 long chains", 196, p.multiProcessorCount, clk, out);
  run<3>("D 98 x IMAD.WIDE.X carry-chained rows", 98, p.multiProcessorCount, clk, out);
  return 0;
}
