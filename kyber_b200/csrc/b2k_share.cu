// b2k_share.cu -- share.RecoverCommit on the device (bn254 G1): Lagrange weights + MSM.
//
// Replaces share.RecoverCommit, /root/reference share/poly.go:449-476: the O(t^2) mod.Int products
// (num *= x_j, den *= x_j - x_i, poly.go:464-470), the t modular inversions (num.Div, :471) and the
// t Point.Mul + Add (:471-472).  The caller (Go: xyCommit, poly.go:418-445) still sorts the shares by index
// and passes the first t (index, point) pairs; x_i = index_i + 1.
#include "msm_host.cuh"
using namespace b2k_host;

namespace b2k {

using Fr254 = Fp<Bn254Fr>;

// lambda_i = prod_{j != i} x_j / (x_j - x_i)  mod r, written as 32-byte big-endian scalars
__global__ void __launch_bounds__(128) k_lagrange_at_zero(uint32_t t, const uint32_t* __restrict__ idx,
                                                          uint8_t* __restrict__ scalars, uint32_t* flags) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= t) return;
  Fr254 num, den, xi, xj, d;
  fp_set_one(num);
  fp_set_one(den);
  fp_set_zero(xi);
  const uint32_t ii = idx[i];
  xi.v[0] = ii + 1u; xi.v[1] = (ii == 0xffffffffu) ? 1u : 0u;
  fp_to_mont(xi, xi);
  for (uint32_t j = 0; j < t; j++) {
    if (j == i) continue;
    const uint32_t jj = idx[j];
    if (jj == ii) atomicOr(flags, 4u);          // duplicate index: denominator would be zero
    fp_set_zero(xj);
    xj.v[0] = jj + 1u; xj.v[1] = (jj == 0xffffffffu) ? 1u : 0u;
    fp_to_mont(xj, xj);
    fp_mul(num, num, xj);
    fp_sub(d, xj, xi);
    fp_mul(den, den, d);
  }
  fp_inv(den, den);
  fp_mul(num, num, den);
  fp_from_mont(num, num);
  uint8_t* out = scalars + 32 * (size_t)i;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    uint8_t* q = out + 4 * (7 - k);
    q[0] = (uint8_t)(num.v[k] >> 24); q[1] = (uint8_t)(num.v[k] >> 16); q[2] = (uint8_t)(num.v[k] >> 8); q[3] = (uint8_t)num.v[k];
  }
}

}  // namespace b2k

extern "C" {

// out = sum_i lambda_i * points[i],  lambda_i from the share indices (x_i = idx_i + 1)
int b2k_bn254_recover_commit(b2k_ctx* ctx, size_t t, const uint32_t* indices /*[t]*/, const uint8_t* points /*[t][64]*/,
                             uint8_t* out /*[64]*/) {
  if (!ctx || !indices || !points || !out || t == 0 || t >= (size_t(1) << 31)) return B2K_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  MsmPlan pl = make_plan(t, ctx->force_c, Bn254G1::SCALAR_BITS);
  size_t extra = pad256(t * 4) + pad256(t * 32) + pad256(t * 64) + 1024;
  int rc = arena_reserve(ctx, msm_scratch_bytes<Bn254G1>(t, pl, ctx->force_L) + extra);
  if (rc) return rc;
  uint32_t* d_idx = arena_take<uint32_t>(ctx, t);
  uint8_t* d_s = arena_take<uint8_t>(ctx, t * 32);
  uint8_t* d_p = arena_take<uint8_t>(ctx, t * 64);
  uint8_t* d_o = arena_take<uint8_t>(ctx, 256);
  cudaStream_t st = ctx->stream;
  CK(cudaMemsetAsync(ctx->d_flags, 0, 4, st));
  CK(cudaMemcpyAsync(d_idx, indices, t * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_p, points, t * 64, cudaMemcpyHostToDevice, st));
  k_lagrange_at_zero<<<(unsigned)((t + 127) / 128), 128, 0, st>>>((uint32_t)t, d_idx, d_s, ctx->d_flags);
  ctx->launches += 1;
  rc = msm_enqueue<Bn254G1>(ctx, t, pl, d_s, d_p, d_o);
  if (rc) return rc;
  CK(cudaMemcpyAsync(out, d_o, 64, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (*ctx->h_flags & 4u) { ctx->err = "duplicate share index"; return B2K_ERR_ARG; }
  return check_flags(ctx);
}

}  // extern "C"
