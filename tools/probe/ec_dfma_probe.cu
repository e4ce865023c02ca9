// build-only probe (instruction mix / registers of the FP64-form mixed XYZZ addition; to be timed on the GPU next round)
#include <cstdio>
#include "ec_dfma.cuh"
__global__ void __launch_bounds__(128, 3) k_dfma_madd_chain(const double* __restrict__ in, double* __restrict__ out, int n) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  dfma::Xyzz acc;
  dfma::Fp one;
  for (int i = 0; i < dfma::L; i++) { acc.X.v[i] = 0; acc.Y.v[i] = 0; acc.ZZ.v[i] = 0; acc.ZZZ.v[i] = 0; one.v[i] = in[i]; }
  int rare = 0;
  for (int k = 0; k < n; k++) {
    dfma::Affine p;
    const double* src = in + dfma::L + ((t * 31 + k) % 4096) * 2 * dfma::L;
    for (int i = 0; i < dfma::L; i++) { p.x.v[i] = src[i]; p.y.v[i] = src[dfma::L + i]; }
    rare += dfma::xyzz_madd<true>(acc, p, one) == 2;
  }
  for (int i = 0; i < dfma::L; i++) out[t * 4 * dfma::L + i] = acc.X.v[i] + acc.Y.v[i] + acc.ZZ.v[i] + acc.ZZZ.v[i] + rare;
}
int main() { printf("build-only\n"); return 0; }
