// K10: the opening phase of create_proof -- polynomial evaluation, Kate division and the inner-product-argument rounds
// (SURVEY.md 8(f) row f2).
//
// Replaces, in halo2_proofs 0.2.0 @ a95945254dcc (Cargo.lock:619-621; reached from /root/reference/src/test_utils.rs:41,96):
//   * arithmetic::eval_polynomial (the ~700 evaluations at x / x*omega^rot written to the transcript)   -> eval_poly_kernel
//   * arithmetic::compute_inner_product                                                                  -> inner_product_kernel
//   * arithmetic::kate_division ((p(X) - p(b)) / (X - b), multiopen)                                     -> kate_* kernels + suffix sums
//   * poly::commitment::prover::create_proof's round body: fold p' and b with the round challenge        -> fold_kernel
//   * arithmetic::parallel_generator_collapse (G'_lo[i] + [u] G'_hi[i], batch-normalised)                -> collapse_* kernels
// The round MSMs <p'_hi, G'_lo>, <p'_lo, G'_hi> run on msm.cu with caller-owned (changing) bases: trp_dev_msm_var.
#include "common.cuh"
#include "ec.cuh"

using namespace ff;
using namespace ec;

namespace {

constexpr int RED_THREADS = 256;
constexpr unsigned CHUNK = 32;   // coefficients per thread in the power-series kernels

template <class PR> __device__ __forceinline__ Fe<PR> fe_shfl_down_(const Fe<PR>& a, unsigned d) {
  Fe<PR> r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = __shfl_down_sync(0xffffffffu, a.v[i], d);
  return r;
}

// sum of `v` over the CTA; valid in thread 0
template <class PR>
__device__ __forceinline__ Fe<PR> block_sum(Fe<PR> v) {
  __shared__ uint4 wsm[2 * (RED_THREADS / 32)];
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fe_add(v, fe_shfl_down_(v, o));
  __syncthreads();
  if (lane == 0) fe_store(wsm + 2 * wid, v);
  __syncthreads();
  if (threadIdx.x == 0) {
    for (unsigned w = 1; w < blockDim.x / 32; ++w) v = fe_add(v, fe_load<PR>(wsm + 2 * w));
  }
  return v;
}

template <class PR> __device__ __forceinline__ Fe<PR> fe_pow_u64(const Fe<PR>& a, uint64_t e) {
  uint32_t l[2] = {(uint32_t)e, (uint32_t)(e >> 32)};
  return fe_pow(a, l, 2);
}

// Evaluation of m polynomials at one point.  A CTA covers RED_THREADS * CHUNK consecutive coefficients; thread t takes the
// coefficients base + t + k * RED_THREADS (coalesced 32-byte loads across the warp), runs Horner in y = x^RED_THREADS over
// them and scales by x^t; thread 0 scales the CTA's sum by x^base.  The powers come from a table made once per call
// (eval_setup_kernel): pw[t] = x^t for t <= RED_THREADS, then x^(b * RED_THREADS * CHUNK) for every CTA b.
// partial[poly][block];  poly = blockIdx.y, addressed as polys + poly * stride or through the pointer table `ptrs`.
template <class PR>
__global__ void eval_setup_kernel(uint4* pw, unsigned blocks, Fe<PR> x) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > RED_THREADS + blocks) return;
  const uint64_t e = i <= RED_THREADS ? (uint64_t)i : (uint64_t)(i - RED_THREADS - 1) * RED_THREADS * CHUNK;
  fe_store(pw + 2 * (size_t)i, e ? fe_pow_u64(x, e) : fe_one<PR>());
}

template <class PR>
__global__ void __launch_bounds__(RED_THREADS) eval_poly_kernel(const uint4* polys, size_t stride, const uint4* const* ptrs, size_t n,
                                                                const uint4* pw, uint4* partial) {
  const uint4* c = ptrs ? ptrs[blockIdx.y] : polys + 2 * (size_t)blockIdx.y * stride;
  const size_t i0 = (size_t)blockIdx.x * (RED_THREADS * CHUNK) + threadIdx.x;
  Fe<PR> acc = fe_zero<PR>();
  if (i0 < n) {
    size_t kmax = (n - 1 - i0) / RED_THREADS;
    if (kmax > CHUNK - 1) kmax = CHUNK - 1;
    const Fe<PR> y = fe_load<PR>(pw + 2 * RED_THREADS);
    acc = fe_load<PR>(c + 2 * (i0 + kmax * RED_THREADS));
    for (size_t k = kmax; k-- > 0;) acc = fe_add(fe_mul(acc, y), fe_load<PR>(c + 2 * (i0 + k * RED_THREADS)));
    if (threadIdx.x) acc = fe_mul(acc, fe_load<PR>(pw + 2 * threadIdx.x));
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) {
    if (blockIdx.x) acc = fe_mul(acc, fe_load<PR>(pw + 2 * (size_t)(RED_THREADS + 1 + blockIdx.x)));
    fe_store(partial + 2 * ((size_t)blockIdx.y * gridDim.x + blockIdx.x), acc);
  }
}

// partial[pair][block] = sum_i a[i] * b[i] over the block's grid-stride slice;  pair = blockIdx.y
template <class PR>
__global__ void __launch_bounds__(RED_THREADS) inner_product_kernel(const uint4* a, size_t a_stride, const uint4* b, size_t b_stride,
                                                                     size_t n, uint4* partial) {
  const uint4* pa = a + 2 * (size_t)blockIdx.y * a_stride;
  const uint4* pb = b + 2 * (size_t)blockIdx.y * b_stride;
  Fe<PR> acc = fe_zero<PR>();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    acc = fe_add(acc, fe_mul(fe_load<PR>(pa + 2 * i), fe_load<PR>(pb + 2 * i)));
  acc = block_sum(acc);
  if (threadIdx.x == 0) fe_store(partial + 2 * ((size_t)blockIdx.y * gridDim.x + blockIdx.x), acc);
}

// out[row] = sum of partial[row][0..count)
template <class PR>
__global__ void __launch_bounds__(RED_THREADS) sum_partials_kernel(const uint4* partial, unsigned count, uint4* out) {
  Fe<PR> acc = fe_zero<PR>();
  for (unsigned i = threadIdx.x; i < count; i += blockDim.x) acc = fe_add(acc, fe_load<PR>(partial + 2 * ((size_t)blockIdx.x * count + i)));
  acc = block_sum(acc);
  if (threadIdx.x == 0) fe_store(out + 2 * (size_t)blockIdx.x, acc);
}

// out[i] = sum_j scal[j] * polys[j][i]: the random linear combinations of multiopen (q_polys = sum_j x_1^e_j * p_j over the ~700
// opened polynomials, poly/multiopen/prover.rs) in one pass over separately allocated inputs
template <class PR>
__global__ void __launch_bounds__(RED_THREADS) lincomb_kernel(const uint4* const* polys, const uint4* scal, unsigned m, size_t n, uint4* out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fe<PR> acc = fe_zero<PR>();
  for (unsigned j = 0; j < m; ++j) acc = fe_add(acc, fe_mul(fe_load<PR>(polys[j] + 2 * i), fe_load<PR>(scal + 2 * (size_t)j)));
  fe_store(out + 2 * i, acc);
}

// a[i] += a[i + half] * u, i < half   (p' with u^-1 and b with u in the IPA round)
template <class PR>
__global__ void fold_kernel(uint4* a, size_t half, Fe<PR> u) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= half) return;
  fe_store(a + 2 * i, fe_add(fe_load<PR>(a + 2 * i), fe_mul(fe_load<PR>(a + 2 * (i + half)), u)));
}

// The IPA round WITHOUT collapsing the generators.  After j rounds halo2's G'_i is sum_t s_t G_{t cur + i}
// (cur = n / 2^j entries alive, s = the 2^j products of the challenges so far), so the round's
//   L_j = <p'_hi, G'_lo> = sum_{t, i < half} p'[half + i] s_t G_{t cur + i},   R_j = <p'_lo, G'_hi> = sum_{t, i < half} p'[i] s_t G_{t cur + half + i}
// are ONE fixed-base MSM each over the ORIGINAL generators (the resident window table): no parallel_generator_collapse (n / 2^(j+1)
// variable-base scalar multiplications per round) and no doubling chain per round.  This kernel writes the two scalar columns
// for the base indices [lo, lo + count): out[0][b - lo], out[1][b - lo].
template <class PR>
__global__ void ipa_round_scalars_kernel(const uint4* p, const uint4* s, unsigned cur_log, size_t lo, size_t count, size_t col_stride, uint4* out) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= count) return;
  const size_t b = lo + idx, cur = (size_t)1 << cur_log, half = cur >> 1;
  const size_t t = b >> cur_log, i = b & (cur - 1);
  const bool low = i < half;
  const Fe<PR> prod = fe_mul(fe_load<PR>(p + 2 * (low ? half + i : i - half)), fe_load_ro<PR>(s + 2 * t));
  const Fe<PR> zero = fe_zero<PR>();
  fe_store(out + 2 * idx, low ? prod : zero);
  fe_store(out + 2 * (col_stride + idx), low ? zero : prod);
}
// s'[2 t] = s[t], s'[2 t + 1] = s[t] * u: G'_next[i] = G'_cur[i] + [u] G'_cur[i + half] in terms of the original generators
template <class PR>
__global__ void ipa_s_double_kernel(const uint4* s, size_t m, Fe<PR> u, uint4* out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  const Fe<PR> v = fe_load<PR>(s + 2 * t);
  fe_store(out + 4 * t, v);
  fe_store(out + 4 * t + 2, fe_mul(v, u));
}

// out[i] = x^i, i < n
template <class PR>
__global__ void powers_kernel(uint4* out, size_t n, Fe<PR> x) {
  const size_t start = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * CHUNK;
  if (start >= n) return;
  Fe<PR> w = fe_pow_u64(x, start);
  for (size_t k = start; k < start + CHUNK && k < n; ++k) { fe_store(out + 2 * k, w); w = fe_mul(w, x); }
}

// kate_division, step 1: t[k] = c[k] * b^k;  step 3: q[i] = S[i] * b^-(i+1), i < n - 1  (S = exclusive suffix sums of t)
template <class PR>
__global__ void kate_scale_kernel(const uint4* in, uint4* out, size_t n, Fe<PR> g, unsigned shift) {
  const size_t start = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * CHUNK;
  if (start >= n) return;
  Fe<PR> w = fe_pow_u64(g, start + shift);
  for (size_t k = start; k < start + CHUNK && k < n; ++k) { fe_store(out + 2 * k, fe_mul(fe_load<PR>(in + 2 * k), w)); w = fe_mul(w, g); }
}
template <class PR>
__global__ void shift_down_kernel(const uint4* in, uint4* out, size_t n_out) {   // b = 0: q[i] = c[i + 1]
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_out) fe_store(out + 2 * i, fe_load<PR>(in + 2 * (i + 1)));
}

// ---- parallel_generator_collapse -----------------------------------------------------------------------------------------
// acc[i] = g_lo[i] + [u] g_hi[i] as XYZZ, den[i] = ZZ * ZZZ (0 for the identity); u given as its 255 canonical bits
struct ScalarBits { uint32_t v[8]; int top; };
template <class BPR>
__global__ void __launch_bounds__(128) collapse_mul_kernel(const uint4* g, size_t half, ScalarBits u, uint4* acc, uint4* den) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= half) return;
  Affine<BPR> lo, hi;
  lo.x = fe_load<BPR>(g + 4 * i); lo.y = fe_load<BPR>(g + 4 * i + 2);
  hi.x = fe_load<BPR>(g + 4 * (i + half)); hi.y = fe_load<BPR>(g + 4 * (i + half) + 2);
  XYZZ<BPR> r = xyzz_identity<BPR>();
  for (int bit = u.top; bit >= 0; --bit) {
    xyzz_dbl(r);
    if ((u.v[bit >> 5] >> (bit & 31)) & 1) xyzz_add_mixed(r, hi);
  }
  xyzz_add_mixed(r, lo);
  uint4* o = acc + 8 * i;
  fe_store(o, r.x); fe_store(o + 2, r.y); fe_store(o + 4, r.zz); fe_store(o + 6, r.zzz);
  fe_store(den + 2 * i, fe_mul(r.zz, r.zzz));
}
// g_lo[i] = (X * ZZZ * inv, Y * ZZ * inv), inv = 1 / (ZZ * ZZZ); identity -> (0, 0)
template <class BPR>
__global__ void __launch_bounds__(128) collapse_norm_kernel(const uint4* acc, const uint4* inv, size_t half, uint4* g) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= half) return;
  const uint4* a = acc + 8 * i;
  Fe<BPR> zz = fe_load<BPR>(a + 4);
  Fe<BPR> x = fe_zero<BPR>(), y = fe_zero<BPR>();
  if (!fe_is_zero(zz)) {
    Fe<BPR> iv = fe_load<BPR>(inv + 2 * i);
    x = fe_mul(fe_load<BPR>(a), fe_mul(iv, fe_load<BPR>(a + 6)));
    y = fe_mul(fe_load<BPR>(a + 2), fe_mul(iv, zz));
  }
  fe_store(g + 4 * i, x); fe_store(g + 4 * i + 2, y);
}

template <class PR> Fe<PR> fe_from_limbs(const uint64_t* l) {
  Fe<PR> r;
  for (int i = 0; i < 4; ++i) { r.v[2 * i] = (uint32_t)l[i]; r.v[2 * i + 1] = (uint32_t)(l[i] >> 32); }
  return r;
}

inline unsigned reduce_blocks(trp_ctx* ctx, size_t n, size_t per_block) {
  size_t b = (n + per_block - 1) / per_block;
  size_t cap = (size_t)ctx->sm_count * 8;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

template <class PR>
int eval_polys_run(trp_ctx* ctx, const void* d_polys, size_t stride, const void* const* d_ptrs, size_t n, size_t m, const uint64_t x[4],
                   void* d_out, void* ws) {
  const unsigned blocks = (unsigned)((n + (size_t)RED_THREADS * CHUNK - 1) / ((size_t)RED_THREADS * CHUNK));
  Fe<PR> fx = fe_from_limbs<PR>(x);
  const size_t mm_max = m < 65535 ? m : 65535;
  uint4* pw = (uint4*)((char*)ws + ws_align((size_t)blocks * mm_max * 32));
  const unsigned entries = RED_THREADS + 1 + blocks;
  eval_setup_kernel<PR><<<(entries + 63) / 64, 64, 0, ctx->stream>>>(pw, blocks, fx);
  TRP_LAUNCHED(ctx);
  for (size_t m0 = 0; m0 < m; m0 += 65535) {
    unsigned mm = (unsigned)(m - m0 < 65535 ? m - m0 : 65535);
    eval_poly_kernel<PR><<<dim3(blocks, mm), RED_THREADS, 0, ctx->stream>>>(d_ptrs ? nullptr : (const uint4*)d_polys + 2 * m0 * stride, stride,
                                                                           d_ptrs ? (const uint4* const*)d_ptrs + m0 : nullptr, n, pw, (uint4*)ws);
    TRP_LAUNCHED(ctx);
    sum_partials_kernel<PR><<<mm, RED_THREADS, 0, ctx->stream>>>((const uint4*)ws, blocks, (uint4*)d_out + 2 * m0);
    TRP_LAUNCHED(ctx);
  }
  return TRP_OK;
}

template <class PR>
int inner_products_run(trp_ctx* ctx, const void* d_a, size_t a_stride, const void* d_b, size_t b_stride, size_t n, size_t m, void* d_out, void* ws) {
  const unsigned blocks = reduce_blocks(ctx, n, (size_t)RED_THREADS * 8);
  for (size_t m0 = 0; m0 < m; m0 += 65535) {
    unsigned mm = (unsigned)(m - m0 < 65535 ? m - m0 : 65535);
    inner_product_kernel<PR><<<dim3(blocks, mm), RED_THREADS, 0, ctx->stream>>>((const uint4*)d_a + 2 * m0 * a_stride, a_stride,
                                                                                  (const uint4*)d_b + 2 * m0 * b_stride, b_stride, n, (uint4*)ws);
    TRP_LAUNCHED(ctx);
    sum_partials_kernel<PR><<<mm, RED_THREADS, 0, ctx->stream>>>((const uint4*)ws, blocks, (uint4*)d_out + 2 * m0);
    TRP_LAUNCHED(ctx);
  }
  return TRP_OK;
}

}  // namespace

int trp_lincomb_impl(trp_ctx* ctx, int field, const void* const* d_ptrs, const void* d_scal, size_t m, size_t n, void* d_out) {
  if (n == 0) return TRP_OK;
  const unsigned blocks = (unsigned)((n + RED_THREADS - 1) / RED_THREADS);
  if (field == 0) lincomb_kernel<FpParams><<<blocks, RED_THREADS, 0, ctx->stream>>>((const uint4* const*)d_ptrs, (const uint4*)d_scal, (unsigned)m, n, (uint4*)d_out);
  else lincomb_kernel<FqParams><<<blocks, RED_THREADS, 0, ctx->stream>>>((const uint4* const*)d_ptrs, (const uint4*)d_scal, (unsigned)m, n, (uint4*)d_out);
  TRP_LAUNCHED(ctx);
  return TRP_OK;
}

// scratch for m simultaneous reductions over n elements
size_t trp_reduce_ws_bytes(trp_ctx* ctx, size_t n, size_t m) {
  size_t b1 = (n + (size_t)RED_THREADS * CHUNK - 1) / ((size_t)RED_THREADS * CHUNK);
  size_t b2 = reduce_blocks(ctx, n, (size_t)RED_THREADS * 8);
  size_t mm = m < 65535 ? m : 65535;
  return ws_align((b1 > b2 ? b1 : b2) * (mm ? mm : 1) * 32) + ws_align((RED_THREADS + 2 + b1) * 32) + 256;   // partials + the power table
}

int trp_eval_polys_impl(trp_ctx* ctx, int field, const void* d_polys, size_t stride, const void* const* d_ptrs, size_t n, size_t m,
                        const uint64_t x[4], void* d_out, void* ws) {
  if (m == 0) return TRP_OK;
  if (n == 0) { TRP_CUDA(ctx, cudaMemsetAsync(d_out, 0, m * 32, ctx->stream)); return TRP_OK; }
  return field == 0 ? eval_polys_run<FpParams>(ctx, d_polys, stride, d_ptrs, n, m, x, d_out, ws)
                    : eval_polys_run<FqParams>(ctx, d_polys, stride, d_ptrs, n, m, x, d_out, ws);
}

int trp_inner_products_impl(trp_ctx* ctx, int field, const void* d_a, size_t a_stride, const void* d_b, size_t b_stride, size_t n, size_t m,
                            void* d_out, void* ws) {
  if (m == 0) return TRP_OK;
  if (n == 0) { TRP_CUDA(ctx, cudaMemsetAsync(d_out, 0, m * 32, ctx->stream)); return TRP_OK; }
  return field == 0 ? inner_products_run<FpParams>(ctx, d_a, a_stride, d_b, b_stride, n, m, d_out, ws)
                    : inner_products_run<FqParams>(ctx, d_a, a_stride, d_b, b_stride, n, m, d_out, ws);
}

int trp_fold_impl(trp_ctx* ctx, int field, void* d_a, size_t half, const uint64_t u[4]) {
  if (half == 0) return TRP_OK;
  unsigned blocks = (unsigned)((half + 255) / 256);
  if (field == 0) fold_kernel<FpParams><<<blocks, 256, 0, ctx->stream>>>((uint4*)d_a, half, fe_from_limbs<FpParams>(u));
  else fold_kernel<FqParams><<<blocks, 256, 0, ctx->stream>>>((uint4*)d_a, half, fe_from_limbs<FqParams>(u));
  TRP_LAUNCHED(ctx);
  return TRP_OK;
}

int trp_ipa_round_scalars_impl(trp_ctx* ctx, int field, const void* d_p, const void* d_s, unsigned cur_log, size_t lo, size_t count,
                                size_t col_stride, void* d_out) {
  if (count == 0) return TRP_OK;
  unsigned blocks = (unsigned)((count + 255) / 256);
  if (field == 0) ipa_round_scalars_kernel<FpParams><<<blocks, 256, 0, ctx->stream>>>((const uint4*)d_p, (const uint4*)d_s, cur_log, lo, count, col_stride, (uint4*)d_out);
  else ipa_round_scalars_kernel<FqParams><<<blocks, 256, 0, ctx->stream>>>((const uint4*)d_p, (const uint4*)d_s, cur_log, lo, count, col_stride, (uint4*)d_out);
  TRP_LAUNCHED(ctx);
  return TRP_OK;
}

int trp_ipa_s_double_impl(trp_ctx* ctx, int field, const void* d_s, size_t m, const uint64_t u[4], void* d_out) {
  if (m == 0) return TRP_OK;
  unsigned blocks = (unsigned)((m + 255) / 256);
  if (field == 0) ipa_s_double_kernel<FpParams><<<blocks, 256, 0, ctx->stream>>>((const uint4*)d_s, m, fe_from_limbs<FpParams>(u), (uint4*)d_out);
  else ipa_s_double_kernel<FqParams><<<blocks, 256, 0, ctx->stream>>>((const uint4*)d_s, m, fe_from_limbs<FqParams>(u), (uint4*)d_out);
  TRP_LAUNCHED(ctx);
  return TRP_OK;
}

int trp_powers_impl(trp_ctx* ctx, int field, const uint64_t x[4], size_t n, void* d_out) {
  if (n == 0) return TRP_OK;
  unsigned blocks = (unsigned)((n + 128 * (size_t)CHUNK - 1) / (128 * (size_t)CHUNK));
  if (field == 0) powers_kernel<FpParams><<<blocks, 128, 0, ctx->stream>>>((uint4*)d_out, n, fe_from_limbs<FpParams>(x));
  else powers_kernel<FqParams><<<blocks, 128, 0, ctx->stream>>>((uint4*)d_out, n, fe_from_limbs<FqParams>(x));
  TRP_LAUNCHED(ctx);
  return TRP_OK;
}

// q (n - 1 coefficients) = (p(X) - p(b)) / (X - b);  ws: n elements + suffix-sum tiles.  b and b_inv Montgomery; b_is_zero
// selects the degenerate case q[i] = c[i + 1].
int trp_suffix_sum_impl(trp_ctx* ctx, int field, void* d_a, size_t n, void* d_tiles);
size_t trp_grand_product_ws_bytes(size_t n_out);
size_t trp_kate_ws_bytes(size_t n) { return ws_align((n ? n : 1) * 32) + trp_grand_product_ws_bytes(n) + 256; }

int trp_kate_division_impl(trp_ctx* ctx, int field, const void* d_coeffs, size_t n, const uint64_t b[4], const uint64_t b_inv[4],
                           int b_is_zero, void* d_q, void* ws) {
  if (n <= 1) return TRP_OK;
  auto run = [&](auto tag) -> int {
    typedef decltype(tag) PR;
    if (b_is_zero) {
      shift_down_kernel<PR><<<(unsigned)((n - 1 + 255) / 256), 256, 0, ctx->stream>>>((const uint4*)d_coeffs, (uint4*)d_q, n - 1);
      TRP_LAUNCHED(ctx);
      return TRP_OK;
    }
    uint4* t = (uint4*)ws;
    void* tiles = (char*)ws + ws_align(n * 32);
    unsigned blocks = (unsigned)((n + 128 * (size_t)CHUNK - 1) / (128 * (size_t)CHUNK));
    kate_scale_kernel<PR><<<blocks, 128, 0, ctx->stream>>>((const uint4*)d_coeffs, t, n, fe_from_limbs<PR>(b), 0);
    TRP_LAUNCHED(ctx);
    TRP_TRY(trp_suffix_sum_impl(ctx, field, t, n, tiles));
    kate_scale_kernel<PR><<<blocks, 128, 0, ctx->stream>>>(t, (uint4*)d_q, n - 1, fe_from_limbs<PR>(b_inv), 1);
    TRP_LAUNCHED(ctx);
    return TRP_OK;
  };
  return field == 0 ? run(FpParams()) : run(FqParams());
}

// g (2 * half affine points) -> g[0..half) = g_lo + [u] g_hi, normalised.  u: CANONICAL scalar limbs.  ws: half * (128 + 32) B.
size_t trp_collapse_ws_bytes(size_t half) { return ws_align(half * 128) + ws_align(half * 32) + 256; }

int trp_generator_collapse_impl(trp_ctx* ctx, void* d_g, size_t half, const uint64_t u_canonical[4], void* ws) {
  if (half == 0) return TRP_OK;
  ScalarBits sb;
  sb.top = -1;
  for (int i = 0; i < 4; ++i) { sb.v[2 * i] = (uint32_t)u_canonical[i]; sb.v[2 * i + 1] = (uint32_t)(u_canonical[i] >> 32); }
  for (int bit = 255; bit >= 0; --bit) if ((sb.v[bit >> 5] >> (bit & 31)) & 1) { sb.top = bit; break; }
  uint4* acc = (uint4*)ws;
  uint4* den = (uint4*)((char*)ws + ws_align(half * 128));
  const int bf = base_field_of(ctx->curve);
  unsigned blocks = (unsigned)((half + 127) / 128);
  if (bf == 0) collapse_mul_kernel<FpParams><<<blocks, 128, 0, ctx->stream>>>((const uint4*)d_g, half, sb, acc, den);
  else collapse_mul_kernel<FqParams><<<blocks, 128, 0, ctx->stream>>>((const uint4*)d_g, half, sb, acc, den);
  TRP_LAUNCHED(ctx);
  TRP_TRY(trp_batch_invert_impl(ctx, bf, den, nullptr, den, half));
  if (bf == 0) collapse_norm_kernel<FpParams><<<blocks, 128, 0, ctx->stream>>>(acc, den, half, (uint4*)d_g);
  else collapse_norm_kernel<FqParams><<<blocks, 128, 0, ctx->stream>>>(acc, den, half, (uint4*)d_g);
  TRP_LAUNCHED(ctx);
  return TRP_OK;
}
