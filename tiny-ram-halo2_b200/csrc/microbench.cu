// Integer-pipe microbenchmarks: establish the roofline denominator for the MSM / NTT kernels
// (MEASURED_PEAKS.json has HBM and bf16 peaks only).  Each kernel issues long runs of one instruction
// class from many resident warps; the result is thread-level instructions per second.  The SASS of every kernel
// was checked with cuobjdump (profiles/int_pipe_r01.md): on sm_100a a 32x32->64 multiply-add costs TWO issue slots
// of the fmaheavy pipe whichever way it is written (IMAD.WIDE.U32[.X] runs at 32 lanes/clk/SM, IMAD lo / IMAD.HI at
// 64 each), so the wide-MAC peak is half the 32-bit IMAD rate: 148 x 32 x 1.965 GHz = 9.3 T MAC/s.
#include "common.cuh"

using namespace ff;

namespace {

constexpr int MB_THREADS = 256;
constexpr int MB_UNROLL = 16;

// kind 0 is mb_wide_carryout below (independent wide MACs with carry-out).  An earlier version of kind 0 multiplied two
// loop-invariant registers; ptxas hoisted the product and the loop became IADD3/IADD3.X pairs (no IMAD in the SASS), so
// the 18.4 T/s it reported was an ALU-pipe number, not an integer-multiply peak.
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
      for (int k = 0; k < 8; ++k) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(acc[(k + 3) & 7]), "r"(y));
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


// kind 9: wide MAC with carry-OUT only (no carry-in), the carry counted into a separate register
// (carry-save accumulation): expect IMAD.WIDE.U32 R, P + IADD3.X cnt.  Counts one op per wide MAC.
__global__ void mb_wide_carryout(uint64_t* out, uint32_t a, uint32_t b, int iters) {
  uint32_t lo[4], hi[4], cnt[4];
  uint32_t x0 = a + threadIdx.x, x1 = x0 * 3 + 1, x2 = x0 * 5 + 2, x3 = x0 * 7 + 3, y = b | 1u;
#pragma unroll
  for (int k = 0; k < 4; ++k) { lo[k] = threadIdx.x * 7 + k; hi[k] = threadIdx.x * 13 + k; cnt[k] = 0; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < MB_UNROLL; ++u) {
      asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;" : "+r"(lo[0]), "+r"(hi[0]), "+r"(cnt[0]) : "r"(x0), "r"(y));
      asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;" : "+r"(lo[1]), "+r"(hi[1]), "+r"(cnt[1]) : "r"(x1), "r"(y));
      asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;" : "+r"(lo[2]), "+r"(hi[2]), "+r"(cnt[2]) : "r"(x2), "r"(y));
      asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;" : "+r"(lo[3]), "+r"(hi[3]), "+r"(cnt[3]) : "r"(x3), "r"(y));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) s ^= lo[k] ^ hi[k] ^ cnt[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// kind 10: independent IMAD.WIDE (no carry) interleaved 1:1 with independent 3-input IADD3: do the two pipes overlap?
// counts BOTH instruction kinds (2 ops per pair)
__global__ void mb_wide_plus_iadd3(uint64_t* out, uint32_t a, uint32_t b, int iters) {
  uint64_t acc[4];
  uint32_t s[4];
  uint32_t x = a + threadIdx.x, y = b | 1u, z = a ^ b;
#pragma unroll
  for (int k = 0; k < 4; ++k) { acc[k] = threadIdx.x * 7 + k; s[k] = threadIdx.x + k; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < MB_UNROLL; ++u) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"((uint32_t)(acc[(k + 1) & 3] >> 32)), "r"(y));
        asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(s[k]) : "r"(s[(k + 1) & 3]), "r"(z));
      }
    }
  }
  uint64_t r = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) r ^= acc[k] ^ s[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// kind 11: independent DFMA (FP64 pipe rate on B200)
__global__ void mb_dfma(uint64_t* out, uint32_t a, uint32_t b, int iters) {
  double acc[8];
  double x = 1.0 + 1e-9 * (a + threadIdx.x), y = 1e-3 * (b | 1u);
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = threadIdx.x * 7 + k;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < MB_UNROLL; ++u) {
#pragma unroll
      for (int k = 0; k < 8; ++k) asm volatile("fma.rn.f64 %0, %1, %0, %2;" : "+d"(acc[k]) : "d"(x), "d"(y));
    }
  }
  double r = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) r += acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint64_t)__double_as_longlong(r);
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
      case 0: mb_wide_carryout<<<blocks, MB_THREADS, 0, ctx->stream>>>(out, 12345u, 67891u, iters); ops_per_thread_iter = 4.0 * MB_UNROLL; break;
      case 1: mb_imad<<<blocks, MB_THREADS, 0, ctx->stream>>>(out, 12345u, 67891u, iters); ops_per_thread_iter = 8.0 * MB_UNROLL; break;
      case 2: mb_iadd_carry<<<blocks, MB_THREADS, 0, ctx->stream>>>(out, 12345u, 67891u, iters); ops_per_thread_iter = 8.0 * MB_UNROLL; break;
      case 3: mb_madc_chain<<<blocks, MB_THREADS, 0, ctx->stream>>>(out, 12345u, 67891u, iters); ops_per_thread_iter = 4.0 * MB_UNROLL; break;
      case 4: mb_fe_mul<FqParams><<<blocks, MB_THREADS, 0, ctx->stream>>>(out, 12345u, 67891u, iters); ops_per_thread_iter = MB_UNROLL; break;
      case 5: mb_fe_addsub<FqParams><<<blocks, MB_THREADS, 0, ctx->stream>>>(out, 12345u, 67891u, iters); ops_per_thread_iter = MB_UNROLL; break;
      case 9: mb_wide_carryout<<<blocks, MB_THREADS, 0, ctx->stream>>>(out, 12345u, 67891u, iters); ops_per_thread_iter = 4.0 * MB_UNROLL; break;
      case 10: mb_wide_plus_iadd3<<<blocks, MB_THREADS, 0, ctx->stream>>>(out, 12345u, 67891u, iters); ops_per_thread_iter = 8.0 * MB_UNROLL; break;
      case 11: mb_dfma<<<blocks, MB_THREADS, 0, ctx->stream>>>(out, 12345u, 67891u, iters); ops_per_thread_iter = 8.0 * MB_UNROLL; break;
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
