// g2_scale_debug.cu -- GPU diagnostic for k_scale_points<Bls381G2> + k_column_sum<Bls381G2>: three ways to compute sum_j s_j P_j.
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include "../../kyber_b200/csrc/msm_host.cuh"
#include "../../kyber_b200/csrc/codec.cuh"
using namespace b2k;
template <class CV>
__global__ void k_run(int t, uint8_t* out /*[4][IN_BYTES]*/, int* oncurve) {
  using F = typename CV::F;
  Affine<F> g; CV::generator(g);
  Jac<F> accA, accB, accC, accD;
  jac_set_inf(accA); jac_set_inf(accB); jac_set_inf(accC); jac_set_inf(accD);
  for (int j = 0; j < t; j++) {
    // P_j = (j + 2) G as an affine point
    Jac<F> pj; jac_from_affine(pj, g);
    for (int i = 0; i < j + 1; i++) jac_madd(pj, pj, g);
    Affine<F> p; jac_to_affine(p, pj);
    Scalar256 k;
    for (int w = 0; w < 8; w++) k.v[w] = 0x9e3779b9u * (uint32_t)(j * 8 + w + 1) + 0x7f4a7c15u;
    k.v[7] &= 0x3fffffffu;
    Jac<F> r4, r1;
    scalar_mul_w4<CV>(r4, k, p, InvBingcd{});
    scalar_mul<CV>(r1, k, p);
    jac_add(accA, accA, r4);                       // A: w4, Jacobian terms, general additions
    jac_add(accB, accB, r1);                       // B: plain double-and-add
    Affine<F> a4; jac_to_affine_bg(a4, r4);
    jac_madd(accC, accC, a4);                      // C: w4, affine terms, mixed additions
    Affine<F> a1; jac_to_affine(a1, r1);
    jac_madd(accD, accD, a1);                      // D: plain, affine terms
    if (j == 0) { oncurve[4] = aff_on_curve<CV>(a4); oncurve[5] = aff_on_curve<CV>(a1); oncurve[6] = f_eq(a4.x, a1.x) && f_eq(a4.y, a1.y); }
  }
  Affine<F> o;
  jac_to_affine(o, accA); CV::store_affine(out, o); oncurve[0] = aff_on_curve<CV>(o);
  jac_to_affine(o, accB); CV::store_affine(out + CV::IN_BYTES, o); oncurve[1] = aff_on_curve<CV>(o);
  jac_to_affine(o, accC); CV::store_affine(out + 2 * CV::IN_BYTES, o); oncurve[2] = aff_on_curve<CV>(o);
  jac_to_affine(o, accD); CV::store_affine(out + 3 * CV::IN_BYTES, o); oncurve[3] = aff_on_curve<CV>(o);
}
template <class CV>
static void run(const char* name, int t) {
  uint8_t* d; int* c; uint8_t h[4 * 192]; int hc[8];
  cudaMalloc(&d, 4 * 192); cudaMalloc(&c, 32); cudaMemset(c, 0, 32);
  k_run<CV><<<1, 1>>>(t, d, c);
  cudaMemcpy(h, d, 4 * CV::IN_BYTES, cudaMemcpyDeviceToHost); cudaMemcpy(hc, c, 32, cudaMemcpyDeviceToHost);
  const int B = CV::IN_BYTES;
  printf("%s t=%d: A==B %d  A==C %d  B==D %d  C==D %d | on curve A %d B %d C %d D %d | term0: w4 on curve %d plain on curve %d equal %d (%s)\n", name, t,
         !memcmp(h, h + B, B), !memcmp(h, h + 2 * B, B), !memcmp(h + B, h + 3 * B, B), !memcmp(h + 2 * B, h + 3 * B, B), hc[0], hc[1], hc[2], hc[3], hc[4], hc[5], hc[6],
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(d); cudaFree(c);
}
int main() {
  for (int t : {1, 2, 3}) { run<Bls381G1>("G1", t); run<Bls381G2>("G2", t); }
  return 0;
}
