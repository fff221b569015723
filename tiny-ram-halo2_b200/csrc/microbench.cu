// Integer-pipe microbenchmarks: establish the roofline denominator for the MSM / NTT kernels
// (MEASURED_PEAKS.json has HBM and bf16 peaks only).  Each kernel issues long runs of one instruction
// class from many resident warps; the result is thread-level instructions per second.
#include "common.cuh"
#include "ff29_experiment.cuh"

using namespace ff;

namespace {

constexpr int MB_THREADS = 256;
constexpr int MB_UNROLL = 16;

// kind 0: independent IMAD.WIDE.U32 (32x32+64 -> 64), 8 accumulators per thread
__global__ void mb_imad_wide(uint64_t* out, uint32_t a, uint32_t b, int iters) {
  uint64_t acc[8];
  uint32_t x = a + threadIdx.x, y = b | 1u;
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = threadIdx.x * 7 + k;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < MB_UNROLL; ++u) {
#pragma unroll
      for (int k = 0; k < 8; ++k) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(x), "r"(y));
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s ^= acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// kind 1: independent 32-bit IMAD
__global__ void mb_imad(uint64_t* out, uint32_t a, uint32_t b, int iters) {
  uint32_t acc[8];
  uint32_t x = a + threadIdx.x, y = b | 1u;
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = threadIdx.x * 7 + k;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < MB_UNROLL; ++u) {
#pragma unroll
      for (int k = 0; k < 8; ++k) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(x), "r"(y));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s ^= acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// kind 2: carry-chained adds (IADD3.X), 8-limb chains
__global__ void mb_iadd_carry(uint64_t* out, uint32_t a, uint32_t b, int iters) {
  uint32_t acc[8];
  uint32_t y = b | 1u;
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = threadIdx.x * 7 + k + a;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < MB_UNROLL; ++u) {
      asm volatile(
          "add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %8;\n\taddc.cc.u32 %2, %2, %8;\n\taddc.cc.u32 %3, %3, %8;\n\t"
          "addc.cc.u32 %4, %4, %8;\n\taddc.cc.u32 %5, %5, %8;\n\taddc.cc.u32 %6, %6, %8;\n\taddc.u32 %7, %7, %8;"
          : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7])
          : "r"(y));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s ^= acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// kind 3: carry-chained wide MACs exactly as in the field multiplier rows (mad.lo.cc / madc.hi.cc pairs);
// counts one op per lo/hi PAIR (= one IMAD.WIDE.X if ptxas fuses them)
__global__ void mb_madc_chain(uint64_t* out, uint32_t a, uint32_t b, int iters) {
  uint32_t acc[8];
  uint32_t x0 = a + threadIdx.x, x1 = x0 * 3 + 1, x2 = x0 * 5 + 2, x3 = x0 * 7 + 3, y = b | 1u;
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = threadIdx.x * 7 + k;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < MB_UNROLL; ++u) {
      asm volatile(
          "mad.lo.cc.u32 %0, %8, %12, %0;\n\tmadc.hi.cc.u32 %1, %8, %12, %1;\n\t"
          "madc.lo.cc.u32 %2, %9, %12, %2;\n\tmadc.hi.cc.u32 %3, %9, %12, %3;\n\t"
          "madc.lo.cc.u32 %4, %10, %12, %4;\n\tmadc.hi.cc.u32 %5, %10, %12, %5;\n\t"
          "madc.lo.cc.u32 %6, %11, %12, %6;\n\tmadc.hi.u32 %7, %11, %12, %7;"
          : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7])
          : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s ^= acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// kind 4: field multiplications, two independent dependency chains per thread
template <class PR>
__global__ void mb_fe_mul(uint64_t* out, uint32_t a, uint32_t b, int iters) {
  Fe<PR> x = fe_one<PR>(), y = fe_one<PR>(), w = fe_one<PR>();
  x.v[0] += threadIdx.x + a; y.v[1] += threadIdx.x * 3 + blockIdx.x + b; w.v[2] += (a ^ b) + threadIdx.x * 5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int u = 0; u < MB_UNROLL / 2; ++u) {
      x = fe_mul(x, w);
      y = fe_mul(y, w);
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s ^= x.v[k] ^ y.v[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// kind 5: field add/sub pairs
template <class PR>
__global__ void mb_fe_addsub(uint64_t* out, uint32_t a, uint32_t b, int iters) {
  Fe<PR> x = fe_one<PR>(), y = fe_one<PR>(), w = fe_one<PR>();
  x.v[0] += threadIdx.x + a; y.v[1] += threadIdx.x * 3 + blockIdx.x + b; w.v[2] += (a ^ b) + threadIdx.x * 5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int u = 0; u < MB_UNROLL / 2; ++u) {
      x = fe_add(x, w);
      y = fe_sub(y, x);
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s ^= x.v[k] ^ y.v[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// kind 6: reduced-radix (9 x 29) field multiplications, two independent chains per thread
template <class PR, int V>
__device__ __forceinline__ ff29::F29<PR> mul29_variant(const ff29::F29<PR>& x, const ff29::F29<PR>& w) {
  if (V == 0) return ff29::mul(x, w);
  if (V == 1) return ff29::mul_cios(x, w);
  return ff29::mul_v3(x, w);
}
template <class PR, int V>
__global__ void mb_mul29(uint64_t* out, uint32_t a, uint32_t b, int iters) {
  ff29::F29<PR> x, y, w;
#pragma unroll
  for (int k = 0; k < 9; ++k) { x.l[k] = (threadIdx.x * 77 + a + k) & ff29::MASK29; y.l[k] = (threadIdx.x * 3 + blockIdx.x + b + k) & ff29::MASK29; w.l[k] = ((a ^ b) + threadIdx.x * 5 + k) & ff29::MASK29; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int u = 0; u < MB_UNROLL / 2; ++u) {
      x = mul29_variant<PR, V>(x, w);
      y = mul29_variant<PR, V>(y, w);
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < 9; ++k) s ^= x.l[k] ^ y.l[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

int trp_microbench_impl(trp_ctx* ctx, int kind, int iters, double* out_gops) {
  if (iters <= 0) iters = 256;
  int blocks = ctx->sm_count * 8;
  TRP_TRY(trp_ws_reserve(ctx, (size_t)blocks * MB_THREADS * 8));
  uint64_t* out = (uint64_t*)ctx->ws;
  cudaEvent_t e0, e1;
  TRP_CUDA(ctx, cudaEventCreate(&e0));
  TRP_CUDA(ctx, cudaEventCreate(&e1));
  double ops_per_thread_iter = 0;
  float best_ms = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    TRP_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    switch (kind) {
      case 0: mb_imad_wide<<<blocks, MB_THREADS, 0, ctx->stream>>>(out, 12345u, 67891u, iters); ops_per_thread_iter = 8.0 * MB_UNROLL; break;
      case 1: mb_imad<<<blocks, MB_THREADS, 0, ctx->stream>>>(out, 12345u, 67891u, iters); ops_per_thread_iter = 8.0 * MB_UNROLL; break;
      case 2: mb_iadd_carry<<<blocks, MB_THREADS, 0, ctx->stream>>>(out, 12345u, 67891u, iters); ops_per_thread_iter = 8.0 * MB_UNROLL; break;
      case 3: mb_madc_chain<<<blocks, MB_THREADS, 0, ctx->stream>>>(out, 12345u, 67891u, iters); ops_per_thread_iter = 4.0 * MB_UNROLL; break;
      case 4: mb_fe_mul<FqParams><<<blocks, MB_THREADS, 0, ctx->stream>>>(out, 12345u, 67891u, iters); ops_per_thread_iter = MB_UNROLL; break;
      case 5: mb_fe_addsub<FqParams><<<blocks, MB_THREADS, 0, ctx->stream>>>(out, 12345u, 67891u, iters); ops_per_thread_iter = MB_UNROLL; break;
      case 6: mb_mul29<ff29::FqP, 0><<<blocks, MB_THREADS, 0, ctx->stream>>>(out, 12345u, 67891u, iters); ops_per_thread_iter = MB_UNROLL; break;
      case 7: mb_mul29<ff29::FqP, 1><<<blocks, MB_THREADS, 0, ctx->stream>>>(out, 12345u, 67891u, iters); ops_per_thread_iter = MB_UNROLL; break;
      case 8: mb_mul29<ff29::FqP, 2><<<blocks, MB_THREADS, 0, ctx->stream>>>(out, 12345u, 67891u, iters); ops_per_thread_iter = MB_UNROLL; break;
      default: TRP_FAIL(ctx, TRP_E_INVALID, "unknown microbenchmark kind %d", kind);
    }
    TRP_LAUNCHED(ctx);
    TRP_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    TRP_CUDA(ctx, cudaEventSynchronize(e1));
    float ms = 0;
    TRP_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best_ms) best_ms = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  double total = (double)blocks * MB_THREADS * (double)iters * ops_per_thread_iter;
  *out_gops = total / (best_ms * 1e-3) / 1e9;
  return TRP_OK;
}
