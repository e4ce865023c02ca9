// Issue-rate probe of the FP64-pipe Montgomery product (tools/probe/fp_dfma.cuh): every thread runs a dependent chain of
// products; products/s for several occupancies, checked against the same code run on the host.  Compare with the IMAD form:
// 3.0e10 products/s (profiles/r01_pipe_probes.txt).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "fp_dfma.cuh"

template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_dfma_chain(const double* __restrict__ in, double* __restrict__ out, int iters) {
  dfma::Fp a, b;
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (int i = 0; i < dfma::L; i++) { a.v[i] = in[(2 * t) * dfma::L + i]; b.v[i] = in[(2 * t + 1) * dfma::L + i]; }
  for (int k = 0; k < iters; k++) { dfma::Fp r; dfma::mont_mul(r, a, b); b = a; a = r; }
  for (int i = 0; i < dfma::L; i++) out[t * dfma::L + i] = a.v[i];
}

static void host_chain(const double* in, double* out, size_t t, int iters) {
  dfma::Fp a, b;
  for (int i = 0; i < dfma::L; i++) { a.v[i] = in[(2 * t) * dfma::L + i]; b.v[i] = in[(2 * t + 1) * dfma::L + i]; }
  for (int k = 0; k < iters; k++) { dfma::Fp r; dfma::mont_mul(r, a, b); b = a; a = r; }
  for (int i = 0; i < dfma::L; i++) out[i] = a.v[i];
}

template <int MINB>
static void run(int sms, int blocks_per_sm, int iters, const double* d_in, double* d_out, const std::vector<double>& h_in) {
  const int grid = sms * blocks_per_sm;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_dfma_chain<MINB><<<grid, 128>>>(d_in, d_out, 16);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_dfma_chain<MINB><<<grid, 128>>>(d_in, d_out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<double> got((size_t)grid * 128 * dfma::L);
  cudaMemcpy(got.data(), d_out, got.size() * 8, cudaMemcpyDeviceToHost);
  int bad = 0;
  const int old = std::fegetround();
  std::fesetround(FE_TOWARDZERO);
  for (size_t t : {size_t(0), size_t(77), (size_t)grid * 128 - 1}) {
    double want[dfma::L];
    host_chain(h_in.data(), want, t, iters);
    for (int i = 0; i < dfma::L; i++) bad += want[i] != got[t * dfma::L + i];
  }
  std::fesetround(old);
  const double prods = (double)grid * 128 * iters;
  printf("blocks/SM %d (launch bound %d): %.3f ms, %.3e products/s, %s: %s\n", blocks_per_sm, MINB, ms, prods / (ms * 1e-3),
         cudaGetErrorString(cudaGetLastError()), bad ? "MISMATCH vs host" : "matches host");
}

int main() {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { printf("no device\n"); return 1; }
  const int sms = prop.multiProcessorCount, max_threads = sms * 8 * 128;
  std::vector<double> h_in((size_t)max_threads * 2 * dfma::L);
  uint64_t s = 88172645463325252ull;
  for (size_t i = 0; i < h_in.size(); i++) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    h_in[i] = (double)((i % dfma::L == dfma::L - 1) ? (s & 0xffff) : (s & dfma::MASK));   // top limb < 2^16: operands below p
  }
  double *d_in, *d_out;
  cudaMalloc(&d_in, h_in.size() * 8);
  cudaMalloc(&d_out, (size_t)max_threads * dfma::L * 8);
  cudaMemcpy(d_in, h_in.data(), h_in.size() * 8, cudaMemcpyHostToDevice);
  printf("%s, %d SMs, clock %d kHz\n", prop.name, sms, prop.clockRate);
  const int iters = 4000;
  run<1>(sms, 1, iters, d_in, d_out, h_in);
  run<1>(sms, 2, iters, d_in, d_out, h_in);
  run<4>(sms, 4, iters, d_in, d_out, h_in);
  run<6>(sms, 6, iters, d_in, d_out, h_in);
  run<8>(sms, 8, iters, d_in, d_out, h_in);
  return 0;
}
