// tree_debug.cu -- GPU diagnostic: block-tree reductions in shared memory vs a sequential loop, for the element types that
// misbehave in k_column_sum<Bls381G2> / k_miller_product<Bn254Pairing> (and the ones that work, as controls).
//   nvcc -O3 -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -o tools/debug/tree_debug tools/debug/tree_debug.cu
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include "../../kyber_b200/csrc/msm_host.cuh"
#include "../../kyber_b200/csrc/codec.cuh"
#include "../../kyber_b200/csrc/bn256.cuh"
#include "../../kyber_b200/csrc/bn_pairing.cuh"
using namespace b2k;

template <class CV>
__global__ void k_make(int n, Jac<typename CV::F>* out) {       // out[j] = (j + 2) G as Jacobian with Z != 1
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  Affine<typename CV::F> g; CV::generator(g);
  Jac<typename CV::F> acc; jac_from_affine(acc, g);
  for (int i = 0; i < j + 1; i++) jac_madd(acc, acc, g);
  jac_dbl(acc, acc);                                            // 2 (j + 2) G, Z != 1
  out[j] = acc;
}
template <class CV>
__global__ void __launch_bounds__(64) k_tree(int t, const Jac<typename CV::F>* terms, uint8_t* out) {
  using J = Jac<typename CV::F>;
  __shared__ J sm[64];
  const int tid = threadIdx.x;
  J acc; jac_set_inf(acc);
  for (int j = tid; j < t; j += 64) { J v = terms[j]; jac_add(acc, acc, v); }
  sm[tid] = acc;
  __syncthreads();
  for (int h = 32; h > 0; h >>= 1) {
    if (tid < h) { J a = sm[tid], b = sm[tid + h]; jac_add(a, a, b); sm[tid] = a; }
    __syncthreads();
  }
  if (tid == 0) { Affine<typename CV::F> a; jac_to_affine(a, sm[0]); CV::store_affine(out, a); }
}
template <class CV>
__global__ void k_seq(int t, const Jac<typename CV::F>* terms, uint8_t* out) {
  using J = Jac<typename CV::F>;
  J acc; jac_set_inf(acc);
  for (int j = 0; j < t; j++) { J v = terms[j]; jac_add(acc, acc, v); }
  Affine<typename CV::F> a; jac_to_affine(a, acc); CV::store_affine(out, a);
}
template <class CV>
static void run(const char* name, int t) {
  using J = Jac<typename CV::F>;
  J* d; uint8_t *o1, *o2; uint8_t h1[256], h2[256];
  cudaMalloc(&d, t * sizeof(J)); cudaMalloc(&o1, 256); cudaMalloc(&o2, 256);
  k_make<CV><<<(t + 31) / 32, 32>>>(t, d);
  k_tree<CV><<<1, 64>>>(t, d, o1);
  k_seq<CV><<<1, 1>>>(t, d, o2);
  cudaMemcpy(h1, o1, CV::IN_BYTES, cudaMemcpyDeviceToHost); cudaMemcpy(h2, o2, CV::IN_BYTES, cudaMemcpyDeviceToHost);
  printf("%s t=%d tree==seq: %s  (%s)\n", name, t, memcmp(h1, h2, CV::IN_BYTES) ? "NO" : "yes", cudaGetErrorString(cudaGetLastError()));
  cudaFree(d); cudaFree(o1); cudaFree(o2);
}
// ---- Fp12 product tree (k_miller_product / k_product_finish shape) vs sequential product --------------------------------------
template <class PC>
__global__ void k_make12(int n, PFp12<PC>* out) {              // out[j] = g^(j + 1), g a fixed non-trivial Fp12 element
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  PFp12<PC> g, acc;
  fp12_set_one(g);
  for (int q = 0; q < PC::FC::N; q++) { g.c0.c1.c0.v[q] = PC::FC::gen_x(q); g.c1.c0.c1.v[q] = PC::FC::gen_y(q); g.c1.c2.c0.v[q] = PC::FC::r2(q); }
  acc = g;
  for (int i = 0; i < j; i++) fp12_mul(acc, acc, g);
  out[j] = acc;
}
template <class PC>
__global__ void __launch_bounds__(32) k_tree12(int t, const PFp12<PC>* terms, PFp12<PC>* out) {
  using F12 = PFp12<PC>;
  extern __shared__ __align__(16) unsigned char smraw[];
  auto* sm = reinterpret_cast<F12*>(smraw);
  const int tid = threadIdx.x;
  F12 f; fp12_set_one(f);
  for (int j = tid; j < t; j += 32) { F12 p = terms[j]; fp12_mul(f, f, p); }
  sm[tid] = f;
  __syncthreads();
  for (int h = 16; h > 0; h >>= 1) {
    if (tid < h) { F12 a = sm[tid], b = sm[tid + h]; fp12_mul(a, a, b); sm[tid] = a; }
    __syncthreads();
  }
  if (tid == 0) out[0] = sm[0];
}
template <class PC>
__global__ void k_seq12(int t, const PFp12<PC>* terms, PFp12<PC>* out) {
  PFp12<PC> f; fp12_set_one(f);
  for (int j = 0; j < t; j++) { PFp12<PC> p = terms[j]; fp12_mul(f, f, p); }
  out[0] = f;
}
template <class PC>
static void run12(const char* name, int t) {
  using F12 = PFp12<PC>;
  F12 *d, *o1, *o2; F12 h1, h2;
  cudaMalloc(&d, t * sizeof(F12)); cudaMalloc(&o1, sizeof(F12)); cudaMalloc(&o2, sizeof(F12));
  k_make12<PC><<<(t + 31) / 32, 32>>>(t, d);
  k_tree12<PC><<<1, 32, 32 * sizeof(F12)>>>(t, d, o1);
  k_seq12<PC><<<1, 1>>>(t, d, o2);
  cudaMemcpy(&h1, o1, sizeof(F12), cudaMemcpyDeviceToHost); cudaMemcpy(&h2, o2, sizeof(F12), cudaMemcpyDeviceToHost);
  printf("%s Fp12 t=%d tree==seq: %s  sizeof=%zu (%s)\n", name, t, memcmp(&h1, &h2, sizeof(F12)) ? "NO" : "yes", sizeof(F12), cudaGetErrorString(cudaGetLastError()));
  cudaFree(d); cudaFree(o1); cudaFree(o2);
}

int main() {
  for (int t : {2, 4, 33}) { run12<Bn254Pair>("bn254", t); run12<Bn256Pair>("bn256", t); }
  for (int t : {1, 2, 6, 17, 70}) { run<Bls381G1>("G1", t); run<Bls381G2>("G2", t); run<Bn254G1>("bn254G1", t); }
  return 0;
}
