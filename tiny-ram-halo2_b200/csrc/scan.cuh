// Exclusive prefix sums of uint32 arrays on the ctx stream (three small kernels, no host synchronisation).
// Shared by the MSM counting sort (msm.cu) and the lookup-permutation radix sort (lookup.cu).
#pragma once
#include "common.cuh"

namespace scan {

// ---- exclusive scan of uint32 (n <= 4096 * 1024) -----------------------------------------------------------
constexpr int SCAN_T = 512, SCAN_ITEMS = 8, SCAN_BLOCK = SCAN_T * SCAN_ITEMS;

static __device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
  __shared__ uint32_t warp_sums[32];
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) warp_sums[wid] = x;
  __syncthreads();
  if (wid == 0) {
    uint32_t s = lane < (blockDim.x >> 5) ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
    warp_sums[lane] = s;
  }
  __syncthreads();
  uint32_t prefix = wid ? warp_sums[wid - 1] : 0;
  if (total) *total = warp_sums[(blockDim.x >> 5) - 1];
  __syncthreads();
  return prefix + x - v;
}

// mode 0: in = counts as is; mode 1: in = ceil((off[i+1]-off[i]) / L) (task counts derived from offsets); mode 2: the number
// of aligned windows [L t, L t + L) that the range [off[i], off[i+1]) meets (the segmented level-1 tasks of msm.cu)
static __device__ __forceinline__ uint32_t scan_input(const uint32_t* in, size_t i, size_t n, int mode, unsigned L) {
  if (i >= n) return 0;
  if (mode == 0) return in[i];
  uint32_t lo = in[i], hi = in[i + 1];
  if (mode == 2) return hi > lo ? (hi - 1) / L - lo / L + 1 : 0;
  return (hi - lo + L - 1) / L;
}

static __global__ void scan_local_kernel(const uint32_t* in, uint32_t* out, uint32_t* block_sums, size_t n, int mode, unsigned L) {
  size_t base = (size_t)blockIdx.x * SCAN_BLOCK + (size_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) { v[k] = scan_input(in, base + k, n, mode, L); sum += v[k]; }
  uint32_t total;
  uint32_t pre = block_exclusive_scan(sum, &total);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) { if (base + k < n) out[base + k] = pre; pre += v[k]; }
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}
static __global__ void scan_sums_kernel(uint32_t* block_sums, unsigned nblocks, uint32_t* grand_total) {
  // single block of SCAN_T threads, nblocks <= SCAN_BLOCK
  size_t base = (size_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) { v[k] = base + k < nblocks ? block_sums[base + k] : 0; sum += v[k]; }
  uint32_t total;
  uint32_t pre = block_exclusive_scan(sum, &total);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) { if (base + k < nblocks) block_sums[base + k] = pre; pre += v[k]; }
  if (threadIdx.x == 0) *grand_total = total;
}
// out[i] += block_sums[block]; optionally mirror into out2 (cursor copy); out[n] = total
static __global__ void scan_add_kernel(uint32_t* out, uint32_t* out2, const uint32_t* block_sums, const uint32_t* grand_total, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    uint32_t v = out[i] + block_sums[i / SCAN_BLOCK];
    out[i] = v;
    if (out2) out2[i] = v;
  } else if (i == n) {
    out[n] = *grand_total;
  }
}


static inline int run_scan(trp_ctx* ctx, const uint32_t* in, uint32_t* out, uint32_t* out2, uint32_t* block_sums, uint32_t* total,
             size_t n, int mode, unsigned L) {
  unsigned nblocks = (unsigned)((n + SCAN_BLOCK - 1) / SCAN_BLOCK);
  if (nblocks > SCAN_BLOCK) TRP_FAIL(ctx, TRP_E_INVALID, "scan of %zu elements exceeds the supported size", n);
  scan_local_kernel<<<nblocks, SCAN_T, 0, ctx->stream>>>(in, out, block_sums, n, mode, L);
  TRP_LAUNCHED(ctx);
  scan_sums_kernel<<<1, SCAN_T, 0, ctx->stream>>>(block_sums, nblocks, total);
  TRP_LAUNCHED(ctx);
  scan_add_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, ctx->stream>>>(out, out2, block_sums, total, n);
  TRP_LAUNCHED(ctx);
  return TRP_OK;
}


}  // namespace scan
