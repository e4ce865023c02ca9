// g2_fine_debug.cu -- GPU diagnostic: where does Jac<Fp2> -> affine go wrong for Z = 1 / accumulated points?
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include "../../kyber_b200/csrc/msm_host.cuh"
#include "../../kyber_b200/csrc/codec.cuh"
using namespace b2k;
__global__ void k(int* res) {
  using CV = Bls381G2;
  using F = CV::F;
  BFp one, zero, t; fp_set_one(one); fp_set_zero(zero);
  fp_sqr(t, zero); res[0] = fp_is_zero(t);
  fp_sqr(t, one); res[1] = fp_eq(t, one);
  fp_inv(t, one); res[2] = fp_eq(t, one);
  F o2, z2; fp2_set_one(o2);
  fp2_inv(z2, o2); res[3] = fp2_eq(z2, o2);
  f_inv_bg(z2, o2); res[4] = fp2_eq(z2, o2);
  Affine<F> g; CV::generator(g);
  Jac<F> j; jac_from_affine(j, g);
  Affine<F> a; jac_to_affine(a, j); res[5] = f_eq(a.x, g.x) && f_eq(a.y, g.y);
  jac_to_affine_bg(a, j); res[6] = f_eq(a.x, g.x) && f_eq(a.y, g.y);
  Jac<F> acc; jac_set_inf(acc);
  res[7] = jac_is_inf(acc);
  jac_add(acc, acc, j); res[8] = f_eq(acc.X, g.x) && f_eq(acc.Y, g.y) && f_eq(acc.Z, o2);
  jac_to_affine(a, acc); res[9] = f_eq(a.x, g.x) && f_eq(a.y, g.y);
  Jac<F> d; jac_dbl(d, j);                       // 2G, Z != 1
  Affine<F> a2, a3; jac_to_affine(a2, d); jac_to_affine_bg(a3, d);
  res[10] = f_eq(a2.x, a3.x) && f_eq(a2.y, a3.y); res[11] = aff_on_curve<CV>(a2); res[12] = aff_on_curve<CV>(a3);
  jac_set_inf(acc); jac_add(acc, acc, d); jac_to_affine(a, acc); res[13] = f_eq(a.x, a2.x) && f_eq(a.y, a2.y);
  jac_set_inf(acc); jac_madd(acc, acc, a2); jac_to_affine(a, acc); res[14] = f_eq(a.x, a2.x) && f_eq(a.y, a2.y);
  res[15] = aff_on_curve<CV>(g);
  uint8_t buf[192]; CV::store_affine(buf, a2); Affine<F> back; CV::load(back, buf); res[16] = f_eq(back.x, a2.x) && f_eq(back.y, a2.y);
}
int main() {
  int* d; int h[32]; cudaMalloc(&d, 128); cudaMemset(d, 0xff, 128);
  k<<<1, 1>>>(d); cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
  const char* names[] = {"sqr(0)==0", "sqr(1)==1", "inv(1)==1", "fp2_inv(1)==1", "fp2_inv_bg(1)==1", "to_affine(G,Z=1)==G", "to_affine_bg(G,Z=1)==G", "set_inf is inf",
                         "inf+G == (G,1)", "to_affine(inf+G)==G", "to_affine(2G)==to_affine_bg(2G)", "2G on curve (plain)", "2G on curve (bg)", "to_affine(inf+2G)==2G",
                         "to_affine(inf+affine 2G)==2G", "G on curve", "store/load round trip"};
  for (int i = 0; i < 17; i++) printf("%-36s %d\n", names[i], h[i]);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
