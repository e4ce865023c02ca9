// kernels.cuh -- __global__ wrappers around the per-thread bodies of msm.cuh (sm_100a).
#pragma once
#include "msm.cuh"

namespace b2k {

// error flag bits written by kernels
enum : uint32_t { FLAG_SCALAR_RANGE = 1u, FLAG_POINT = 2u };

// ---- operand conversion: wire bytes -> Montgomery affine (AoS, one point = 2*N limbs) -----------
template <class CV>
__global__ void __launch_bounds__(256) k_load_points(size_t n, const uint8_t* __restrict__ wire,
                                                     Affine<typename CV::F>* __restrict__ pts) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<typename CV::F> p;
  CV::load(p, wire + (size_t)CV::IN_BYTES * i);
  pts[i] = p;
}

// ---- independent scalar multiplications ----------------------------------------------------------
template <class CV, bool AFFINE_OUT>
__global__ void __launch_bounds__(128) k_mul_batch(size_t n, const uint8_t* __restrict__ scalars,
                                                   const uint8_t* __restrict__ wire, uint8_t* __restrict__ out,
                                                   uint32_t* flags) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Scalar256 k;
  scalar_load_be(k, scalars + 32 * i);
  if (!scalar_in_range<typename CV::ScalarField>(k)) atomicOr(flags, FLAG_SCALAR_RANGE);
  Affine<typename CV::F> p;
  CV::load(p, wire + (size_t)CV::IN_BYTES * i);
  Jac<typename CV::F> r;
  scalar_mul<CV>(r, k, p);
  Affine<typename CV::F> a;
  jac_to_affine(a, r);
  if (AFFINE_OUT) CV::store_affine(out + (size_t)CV::IN_BYTES * i, a);
  else CV::store(out + (size_t)CV::OUT_BYTES * i, a);
}

// ---- MSM stage 1: digits + histogram -------------------------------------------------------------
template <class CV>
__global__ void __launch_bounds__(256) k_msm_count(size_t n, const uint8_t* __restrict__ scalars, MsmPlan pl,
                                                   uint32_t* __restrict__ counts, uint32_t* flags) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Scalar256 s;
  scalar_load_be(s, scalars + 32 * i);
  if (!scalar_in_range<typename CV::ScalarField>(s)) atomicOr(flags, FLAG_SCALAR_RANGE);
  uint32_t sp[9];
  msm_recode(sp, s, pl.K);
  for (int w = 0; w < pl.W; w++) {
    int d = msm_digit(sp, pl.c, w);
    if (d) atomicAdd(&counts[(size_t)w * pl.nb + (d < 0 ? -d : d) - 1], 1u);
  }
}

// ---- MSM stage 2: exclusive scan of the histogram (single block) ---------------------------------
__global__ void __launch_bounds__(1024) k_msm_scan(size_t total, const uint32_t* __restrict__ counts,
                                                   uint32_t* __restrict__ offs, uint32_t* __restrict__ cursor) {
  __shared__ uint32_t sm[1024];
  int tid = threadIdx.x;
  size_t per = (total + 1023) / 1024;
  size_t b = (size_t)tid * per, e = b + per;
  if (b > total) b = total;
  if (e > total) e = total;
  uint32_t s = 0;
  for (size_t i = b; i < e; i++) s += counts[i];
  sm[tid] = s;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {   // Hillis-Steele inclusive scan
    uint32_t v = (tid >= d) ? sm[tid - d] : 0;
    __syncthreads();
    sm[tid] += v;
    __syncthreads();
  }
  uint32_t run = sm[tid] - s;            // exclusive prefix of this thread's slice
  for (size_t i = b; i < e; i++) {
    offs[i] = run;
    cursor[i] = run;
    run += counts[i];
  }
  if (tid == 1023) offs[total] = sm[1023];
}

// ---- MSM stage 3: counting-sort scatter ----------------------------------------------------------
template <class CV>
__global__ void __launch_bounds__(256) k_msm_scatter(size_t n, const uint8_t* __restrict__ scalars, MsmPlan pl,
                                                     uint32_t* __restrict__ cursor, uint32_t* __restrict__ entries) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Scalar256 s;
  scalar_load_be(s, scalars + 32 * i);
  uint32_t sp[9];
  msm_recode(sp, s, pl.K);
  for (int w = 0; w < pl.W; w++) {
    int d = msm_digit(sp, pl.c, w);
    if (d) {
      uint32_t pos = atomicAdd(&cursor[(size_t)w * pl.nb + (d < 0 ? -d : d) - 1], 1u);
      entries[pos] = (uint32_t)i | (d < 0 ? 0x80000000u : 0u);
    }
  }
}

// ---- MSM stage 4: bucket accumulate (THE hot kernel) ---------------------------------------------
template <class CV>
__global__ void __launch_bounds__(128) k_msm_accumulate(size_t total, const Affine<typename CV::F>* __restrict__ pts,
                                                        const uint32_t* __restrict__ offs,
                                                        const uint32_t* __restrict__ entries,
                                                        Xyzz<typename CV::F>* __restrict__ buckets) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  Xyzz<typename CV::F> acc;
  msm_accumulate_bucket<CV>(acc, pts, entries, offs[g], offs[g + 1]);
  buckets[g] = acc;
}

// ---- MSM stage 5: per-chunk running sums ---------------------------------------------------------
template <class CV>
__global__ void __launch_bounds__(128) k_msm_reduce_chunks(MsmPlan pl, const Xyzz<typename CV::F>* __restrict__ buckets,
                                                           Xyzz<typename CV::F>* __restrict__ partials) {
  int T = pl.nb / pl.m;
  size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (size_t)pl.W * T) return;
  int w = (int)(id / T), t = (int)(id % T);
  Xyzz<typename CV::F> out;
  msm_reduce_chunk<CV>(out, buckets + (size_t)w * pl.nb, t, pl.m);
  partials[id] = out;
}

// ---- MSM stage 6: per-window sum of the partials (one block per window) --------------------------
template <class CV>
__global__ void __launch_bounds__(128) k_msm_window_sum(int T, const Xyzz<typename CV::F>* __restrict__ partials,
                                                        Xyzz<typename CV::F>* __restrict__ wsum) {
  using X = Xyzz<typename CV::F>;
  __shared__ X sm[128];
  int tid = threadIdx.x, w = blockIdx.x;
  X acc;
  xyzz_set_inf(acc);
  for (int t = tid; t < T; t += 128) {
    X p = partials[(size_t)w * T + t];
    xyzz_add(acc, acc, p);
  }
  sm[tid] = acc;
  __syncthreads();
  for (int s = 64; s > 0; s >>= 1) {
    if (tid < s) {
      X a = sm[tid], b = sm[tid + s];
      xyzz_add(a, a, b);
      sm[tid] = a;
    }
    __syncthreads();
  }
  if (tid == 0) wsum[w] = sm[0];
}

// ---- MSM stage 7: Horner over windows, affine, wire bytes ------------------------------------------
template <class CV>
__global__ void k_msm_final(MsmPlan pl, const Xyzz<typename CV::F>* __restrict__ wsum, uint8_t* __restrict__ out,
                            int affine_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  Xyzz<typename CV::F> r;
  msm_horner<CV>(r, wsum, pl.W, pl.c);
  Affine<typename CV::F> a;
  xyzz_to_affine(a, r);
  if (affine_out) CV::store_affine(out, a);
  else CV::store(out, a);
}

}  // namespace b2k
