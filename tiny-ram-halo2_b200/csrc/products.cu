// K8: batch inversion and grand products -- the field work that sits BETWEEN the hot kernels inside create_proof
// (SURVEY.md 8(f) row f1).
//
// Replaces, in halo2_proofs 0.2.0 @ a95945254dcc (Cargo.lock:619-621; reached from /root/reference/src/test_utils.rs:41,96):
//   * ff::BatchInvert::batch_invert                                   -> batch_invert_kernel
//   * plonk::permutation::prover::Argument::commit  (the Z columns of the 188 equality-enabled columns the reference
//     declares at /root/reference/src/circuits/tables/prog.rs:151-152)  -> perm_terms_kernel + batch inversion + grand product
//   * plonk::lookup::prover::Permuted::commit_product (31 lookups: /root/reference/src/circuits/tables/even_bits.rs:158-170,
//     out_table.rs:33-74, shift.rs:142-165, circuits/mod.rs:52-57)     -> lookup_terms_kernel + the same two steps
// Left on the CPU these force every column across PCIe twice per proof; here the columns never leave HBM.
//
// Batch inversion: a CTA owns 2048 consecutive elements (256 threads x 8).  Each thread builds the running product of its
// non-zero elements (prefixes parked in shared memory), the 256 thread products are combined with warp-shuffle product
// scans, ONE warp inverts the CTA product (a 255-bit exponentiation, ~380 multiplications, while the other warps wait at
// the barrier and cost no issue slots) and every thread walks back over its elements: 3 multiplications per element plus
// ~30 per thread.  Zero elements stay zero, as in ff::BatchInvert.
// Grand product: three-kernel exclusive product scan with the same CTA shape (tile products -> scan of tile products
// seeded with the initial value -> per-tile walk), ~5 multiplications per element.
#include "common.cuh"

#include <vector>

using namespace ff;

namespace {

constexpr int PB_THREADS = 256, PB_K = 8, PB_TILE = PB_THREADS * PB_K, PB_WARPS = PB_THREADS / 32;
constexpr unsigned MAX_PERM_COLS = 16;

template <class PR> __device__ __forceinline__ Fe<PR> fe_shfl_up(const Fe<PR>& a, unsigned d) {
  Fe<PR> r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = __shfl_up_sync(0xffffffffu, a.v[i], d);
  return r;
}
template <class PR> __device__ __forceinline__ Fe<PR> fe_shfl_down(const Fe<PR>& a, unsigned d) {
  Fe<PR> r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = __shfl_down_sync(0xffffffffu, a.v[i], d);
  return r;
}
template <class PR> __device__ __forceinline__ Fe<PR> fe_shfl(const Fe<PR>& a, unsigned lane) {
  Fe<PR> r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = __shfl_sync(0xffffffffu, a.v[i], lane);
  return r;
}

// Products over the CTA's 256 thread values p: returns the product of the values of all LOWER threads (exclusive prefix);
// optionally also the product of all HIGHER threads and the CTA total.  wsm: PB_WARPS field elements of shared memory.
// ADD = false: the monoid is (field, *, 1); ADD = true: (field, +, 0) -- prefix sums for kate_division
template <class PR, bool ADD> __device__ __forceinline__ Fe<PR> op_identity() { return ADD ? fe_zero<PR>() : fe_one<PR>(); }
template <class PR, bool ADD> __device__ __forceinline__ Fe<PR> op_apply(const Fe<PR>& a, const Fe<PR>& b) { return ADD ? fe_add(a, b) : fe_mul(a, b); }

template <class PR, bool ADD = false>
__device__ __forceinline__ Fe<PR> block_exclusive_products(const Fe<PR>& p, uint4* wsm, Fe<PR>* suffix, Fe<PR>* total) {
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  Fe<PR> inc = p;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { Fe<PR> y = fe_shfl_up(inc, o); if (lane >= (unsigned)o) inc = op_apply<PR, ADD>(inc, y); }
  Fe<PR> exc = fe_shfl_up(inc, 1);
  if (lane == 0) exc = op_identity<PR, ADD>();
  Fe<PR> sexc = op_identity<PR, ADD>();
  if (suffix) {
    Fe<PR> sinc = p;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { Fe<PR> y = fe_shfl_down(sinc, o); if (lane + o < 32) sinc = op_apply<PR, ADD>(sinc, y); }
    sexc = fe_shfl_down(sinc, 1);
    if (lane == 31) sexc = op_identity<PR, ADD>();
  }
  __syncthreads();                      // wsm may still be read by a previous call
  if (lane == 31) fe_store(wsm + 2 * wid, inc);
  __syncthreads();
  Fe<PR> before = op_identity<PR, ADD>(), after = op_identity<PR, ADD>(), all = op_identity<PR, ADD>();
  for (unsigned w = 0; w < (unsigned)PB_WARPS; ++w) {
    Fe<PR> t = fe_load<PR>(wsm + 2 * w);
    if (w < wid) before = op_apply<PR, ADD>(before, t);
    if (suffix && w > wid) after = op_apply<PR, ADD>(after, t);
    if (total) all = op_apply<PR, ADD>(all, t);
  }
  if (suffix) *suffix = op_apply<PR, ADD>(sexc, after);
  if (total) *total = all;
  return op_apply<PR, ADD>(exc, before);
}

// out[i] = mul[i] / a[i]  (mul == nullptr: 1 / a[i]);  a[i] == 0 -> out[i] = 0 (ff::BatchInvert skips zeros)
template <class PR>
__global__ void __launch_bounds__(PB_THREADS) batch_invert_kernel(const uint4* a, const uint4* mul, uint4* out, size_t n) {
  extern __shared__ uint4 smem[];
  uint4* pre0 = smem;                       // prefix products, two 16-byte planes indexed [k][thread]
  uint4* pre1 = smem + PB_TILE;
  uint4* wsm = smem + 2 * PB_TILE;          // PB_WARPS warp products + 1 slot for the inverse of the CTA product
  const unsigned tid = threadIdx.x;
  const size_t base = (size_t)blockIdx.x * PB_TILE + (size_t)tid * PB_K;
  Fe<PR> acc = fe_one<PR>();
#pragma unroll 1
  for (int k = 0; k < PB_K; ++k) {
    pre0[k * PB_THREADS + tid] = make_uint4(acc.v[0], acc.v[1], acc.v[2], acc.v[3]);
    pre1[k * PB_THREADS + tid] = make_uint4(acc.v[4], acc.v[5], acc.v[6], acc.v[7]);
    if (base + k < n) {
      Fe<PR> v = fe_load<PR>(a + 2 * (base + k));
      if (!fe_is_zero(v)) acc = fe_mul(acc, v);
    }
  }
  Fe<PR> suffix, total;
  Fe<PR> prefix = block_exclusive_products(acc, wsm, &suffix, &total);
  if (tid < 32) {
    Fe<PR> ti = fe_inv(total);              // never zero: zero elements were skipped
    if (tid == 0) fe_store(wsm + 2 * PB_WARPS, ti);
  }
  __syncthreads();
  Fe<PR> inv = fe_mul(fe_load<PR>(wsm + 2 * PB_WARPS), fe_mul(prefix, suffix));   // 1 / (this thread's product)
#pragma unroll 1
  for (int k = PB_K - 1; k >= 0; --k) {
    if (base + k >= n) continue;
    Fe<PR> v = fe_load<PR>(a + 2 * (base + k));
    Fe<PR> o = fe_zero<PR>();
    if (!fe_is_zero(v)) {
      uint4 lo = pre0[k * PB_THREADS + tid], hi = pre1[k * PB_THREADS + tid];
      Fe<PR> pr;
      pr.v[0] = lo.x; pr.v[1] = lo.y; pr.v[2] = lo.z; pr.v[3] = lo.w; pr.v[4] = hi.x; pr.v[5] = hi.y; pr.v[6] = hi.z; pr.v[7] = hi.w;
      o = fe_mul(inv, pr);
      inv = fe_mul(inv, v);
      if (mul) o = fe_mul(o, fe_load<PR>(mul + 2 * (base + k)));
    }
    fe_store(out + 2 * (base + k), o);
  }
}

// ---- grand product / prefix sum: z[0] = init, z[i] = z[i-1] (op) v[i-1] for i < n_out (v[j] = identity for j >= n_in) ----
// REV: logical index i lives at physical index (len - 1 - i) of both arrays, i.e. the scan runs from the END (suffix sums).
template <bool REV> __device__ __forceinline__ size_t phys(size_t i, size_t len) { return REV ? len - 1 - i : i; }

template <class PR, bool ADD, bool REV>
__global__ void __launch_bounds__(PB_THREADS) gp_tile_products_kernel(const uint4* v, size_t n_in, uint4* tile_prod) {
  __shared__ uint4 wsm[2 * PB_WARPS];
  const size_t base = (size_t)blockIdx.x * PB_TILE + (size_t)threadIdx.x * PB_K;
  Fe<PR> acc = op_identity<PR, ADD>();
#pragma unroll 1
  for (int k = 0; k < PB_K; ++k)
    if (base + k < n_in) acc = op_apply<PR, ADD>(acc, fe_load<PR>(v + 2 * phys<REV>(base + k, n_in)));
  Fe<PR> total;
  block_exclusive_products<PR, ADD>(acc, wsm, nullptr, &total);
  if (threadIdx.x == 0) fe_store(tile_prod + 2 * (size_t)blockIdx.x, total);
}

// single CTA: tile_prod[b] <- init (op) (op)_{b' < b} tile_prod[b']
template <class PR, bool ADD>
__global__ void __launch_bounds__(PB_THREADS) gp_scan_tiles_kernel(uint4* tile_prod, unsigned ntiles, const uint4* init) {
  __shared__ uint4 wsm[2 * PB_WARPS];
  const unsigned per = (ntiles + PB_THREADS - 1) / PB_THREADS;
  const unsigned lo = threadIdx.x * per, hi = min(lo + per, ntiles);
  Fe<PR> acc = op_identity<PR, ADD>();
  for (unsigned b = lo; b < hi; ++b) acc = op_apply<PR, ADD>(acc, fe_load<PR>(tile_prod + 2 * (size_t)b));
  Fe<PR> run = block_exclusive_products<PR, ADD>(acc, wsm, nullptr, nullptr);
  if (init) run = op_apply<PR, ADD>(run, fe_load<PR>(init));
  __syncthreads();
  for (unsigned b = lo; b < hi; ++b) {
    Fe<PR> t = fe_load<PR>(tile_prod + 2 * (size_t)b);
    fe_store(tile_prod + 2 * (size_t)b, run);
    run = op_apply<PR, ADD>(run, t);
  }
}

template <class PR, bool ADD, bool REV>
__global__ void __launch_bounds__(PB_THREADS) gp_apply_kernel(const uint4* v, size_t n_in, const uint4* tile_prefix, uint4* z, size_t n_out) {
  __shared__ uint4 wsm[2 * PB_WARPS];
  const size_t base = (size_t)blockIdx.x * PB_TILE + (size_t)threadIdx.x * PB_K;
  Fe<PR> acc = op_identity<PR, ADD>();
#pragma unroll 1
  for (int k = 0; k < PB_K; ++k)
    if (base + k < n_in) acc = op_apply<PR, ADD>(acc, fe_load<PR>(v + 2 * phys<REV>(base + k, n_in)));
  Fe<PR> run = block_exclusive_products<PR, ADD>(acc, wsm, nullptr, nullptr);
  run = op_apply<PR, ADD>(run, fe_load<PR>(tile_prefix + 2 * (size_t)blockIdx.x));
#pragma unroll 1
  for (int k = 0; k < PB_K; ++k) {
    if (base + k >= n_out) break;
    // read before z[i] is written: v may alias z (same logical index <-> same physical index when n_in == n_out)
    Fe<PR> cur = base + k < n_in ? fe_load<PR>(v + 2 * phys<REV>(base + k, n_in)) : op_identity<PR, ADD>();
    fe_store(z + 2 * phys<REV>(base + k, n_out), run);
    run = op_apply<PR, ADD>(run, cur);
  }
}

// ---- permutation argument terms -------------------------------------------------------------------------------------------
struct PermCols {
  const uint4* val[MAX_PERM_COLS];
  const uint4* sig[MAX_PERM_COLS];
  unsigned m;
};
// den[i] = prod_c (beta * sigma_c[i] + gamma + v_c[i]);  num[i] = prod_c (dbeta_c * omega^i + gamma + v_c[i]),
// dbeta_c = delta^(first column index + c) * beta   (permutation/prover.rs: "deltaomega * beta + gamma + value")
template <class PR>
__global__ void __launch_bounds__(128) perm_terms_kernel(PermCols cols, const uint4* consts /* beta, gamma, dbeta[m] */,
                                                        const uint4* tw /* omega^i, i < n/2 */, unsigned log_n, uint4* num, uint4* den) {
  const size_t n = (size_t)1 << log_n;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Fe<PR> beta = fe_load_ro<PR>(consts), gamma = fe_load_ro<PR>(consts + 2);
  Fe<PR> w;
  if (log_n == 0) w = fe_one<PR>();
  else {
    const size_t half = n >> 1;
    w = fe_load_ro<PR>(tw + 2 * (i & (half - 1)));
    if (i & half) w = fe_neg(w);
  }
  Fe<PR> nu = fe_one<PR>(), de = fe_one<PR>();
  for (unsigned c = 0; c < cols.m; ++c) {
    Fe<PR> v = fe_add(fe_load<PR>(cols.val[c] + 2 * i), gamma);
    Fe<PR> d = fe_add(fe_mul(beta, fe_load<PR>(cols.sig[c] + 2 * i)), v);
    Fe<PR> u = fe_add(fe_mul(fe_load_ro<PR>(consts + 4 + 2 * c), w), v);
    if (c == 0) { nu = u; de = d; } else { nu = fe_mul(nu, u); de = fe_mul(de, d); }
  }
  fe_store(num + 2 * i, nu);
  fe_store(den + 2 * i, de);
}

// lookup product terms (lookup/prover.rs commit_product): num = (a + beta)(s + gamma), den = (a' + beta)(s' + gamma)
template <class PR>
__global__ void __launch_bounds__(128) lookup_terms_kernel(const uint4* a, const uint4* s, const uint4* ap, const uint4* sp,
                                                          const uint4* consts /* beta, gamma */, size_t n, uint4* num, uint4* den) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Fe<PR> beta = fe_load_ro<PR>(consts), gamma = fe_load_ro<PR>(consts + 2);
  fe_store(num + 2 * i, fe_mul(fe_add(fe_load<PR>(a + 2 * i), beta), fe_add(fe_load<PR>(s + 2 * i), gamma)));
  fe_store(den + 2 * i, fe_mul(fe_add(fe_load<PR>(ap + 2 * i), beta), fe_add(fe_load<PR>(sp + 2 * i), gamma)));
}

constexpr size_t BI_SMEM = (size_t)(2 * PB_TILE + 2 * PB_WARPS + 2) * sizeof(uint4);

template <class PR>
int batch_invert_run(trp_ctx* ctx, const void* d_a, const void* d_mul, void* d_out, size_t n) {
  if (n == 0) return TRP_OK;
  TRP_CUDA(ctx, cudaFuncSetAttribute(batch_invert_kernel<PR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BI_SMEM));
  unsigned blocks = (unsigned)((n + PB_TILE - 1) / PB_TILE);
  ProfScope ps(ctx, PROF_PRODUCTS);
  batch_invert_kernel<PR><<<blocks, PB_THREADS, BI_SMEM, ctx->stream>>>((const uint4*)d_a, (const uint4*)d_mul, (uint4*)d_out, n);
  TRP_LAUNCHED(ctx);
  return TRP_OK;
}

template <class PR>
int grand_product_run(trp_ctx* ctx, const void* d_v, size_t n_in, const void* d_init, void* d_z, size_t n_out, void* d_tiles) {
  if (n_out == 0) return TRP_OK;
  unsigned ntiles = (unsigned)((n_out + PB_TILE - 1) / PB_TILE);
  ProfScope ps(ctx, PROF_PRODUCTS);
  gp_tile_products_kernel<PR, false, false><<<ntiles, PB_THREADS, 0, ctx->stream>>>((const uint4*)d_v, n_in, (uint4*)d_tiles);
  TRP_LAUNCHED(ctx);
  gp_scan_tiles_kernel<PR, false><<<1, PB_THREADS, 0, ctx->stream>>>((uint4*)d_tiles, ntiles, (const uint4*)d_init);
  TRP_LAUNCHED(ctx);
  gp_apply_kernel<PR, false, false><<<ntiles, PB_THREADS, 0, ctx->stream>>>((const uint4*)d_v, n_in, (const uint4*)d_tiles, (uint4*)d_z, n_out);
  TRP_LAUNCHED(ctx);
  return TRP_OK;
}

// in place: a[i] <- sum_{k > i} a[k]   (exclusive suffix sums; the middle step of kate_division)
template <class PR>
int suffix_sum_run(trp_ctx* ctx, void* d_a, size_t n, void* d_tiles) {
  if (n == 0) return TRP_OK;
  unsigned ntiles = (unsigned)((n + PB_TILE - 1) / PB_TILE);
  gp_tile_products_kernel<PR, true, true><<<ntiles, PB_THREADS, 0, ctx->stream>>>((const uint4*)d_a, n, (uint4*)d_tiles);
  TRP_LAUNCHED(ctx);
  gp_scan_tiles_kernel<PR, true><<<1, PB_THREADS, 0, ctx->stream>>>((uint4*)d_tiles, ntiles, nullptr);
  TRP_LAUNCHED(ctx);
  gp_apply_kernel<PR, true, true><<<ntiles, PB_THREADS, 0, ctx->stream>>>((const uint4*)d_a, n, (const uint4*)d_tiles, (uint4*)d_a, n);
  TRP_LAUNCHED(ctx);
  return TRP_OK;
}

inline size_t gp_tiles_bytes(size_t n_out) { return ws_align(((n_out + PB_TILE - 1) / PB_TILE + 1) * 32); }

}  // namespace

int trp_batch_invert_impl(trp_ctx* ctx, int field, const void* d_a, const void* d_mul, void* d_out, size_t n) {
  return field == 0 ? batch_invert_run<FpParams>(ctx, d_a, d_mul, d_out, n) : batch_invert_run<FqParams>(ctx, d_a, d_mul, d_out, n);
}

size_t trp_grand_product_ws_bytes(size_t n_out) { return gp_tiles_bytes(n_out); }

int trp_grand_product_impl(trp_ctx* ctx, int field, const void* d_v, size_t n_in, const void* d_init, void* d_z, size_t n_out, void* d_tiles) {
  return field == 0 ? grand_product_run<FpParams>(ctx, d_v, n_in, d_init, d_z, n_out, d_tiles)
                    : grand_product_run<FqParams>(ctx, d_v, n_in, d_init, d_z, n_out, d_tiles);
}

int trp_suffix_sum_impl(trp_ctx* ctx, int field, void* d_a, size_t n, void* d_tiles) {
  return field == 0 ? suffix_sum_run<FpParams>(ctx, d_a, n, d_tiles) : suffix_sum_run<FqParams>(ctx, d_a, n, d_tiles);
}

size_t trp_product_ws_bytes(size_t n) { return 2 * ws_align(n * 32) + gp_tiles_bytes(n) + ws_align((2 + MAX_PERM_COLS) * 32); }

// z (n = 2^k values) of one chunk of the permutation argument; ws holds num | den | tiles | consts
int trp_permutation_product_impl(trp_domain* d, const uint64_t* const* d_values, const uint64_t* const* d_sigmas, size_t m,
                                 const uint64_t* consts_host /* beta, gamma, dbeta[m]: (2 + m) x 4 */, const void* d_last_z,
                                 void* d_z, void* ws) {
  trp_ctx* ctx = d->ctx;
  if (m == 0 || m > MAX_PERM_COLS) TRP_FAIL(ctx, TRP_E_INVALID, "a permutation chunk holds 1..%u columns (got %zu)", MAX_PERM_COLS, m);
  const size_t n = (size_t)1 << d->k;
  char* p = (char*)ws;
  void* num = p; p += ws_align(n * 32);
  void* den = p; p += ws_align(n * 32);
  void* tiles = p; p += gp_tiles_bytes(n);
  void* dconsts = p;
  TRP_CUDA(ctx, cudaMemcpyAsync(dconsts, consts_host, (2 + m) * 32, cudaMemcpyHostToDevice, ctx->stream));
  PermCols pc;
  pc.m = (unsigned)m;
  for (size_t c = 0; c < m; ++c) {
    if (!d_values[c] || !d_sigmas[c]) TRP_FAIL(ctx, TRP_E_INVALID, "NULL column pointer in permutation chunk");
    pc.val[c] = (const uint4*)d_values[c]; pc.sig[c] = (const uint4*)d_sigmas[c];
  }
  const void* tw = nullptr;
  TRP_TRY(trp_get_powers(ctx, d->field, d->k, d->omega, &tw));
  unsigned blocks = (unsigned)((n + 127) / 128);
  {
    ProfScope ps(ctx, PROF_PRODUCTS);
    if (d->field == 0) perm_terms_kernel<FpParams><<<blocks, 128, 0, ctx->stream>>>(pc, (const uint4*)dconsts, (const uint4*)tw, d->k, (uint4*)num, (uint4*)den);
    else perm_terms_kernel<FqParams><<<blocks, 128, 0, ctx->stream>>>(pc, (const uint4*)dconsts, (const uint4*)tw, d->k, (uint4*)num, (uint4*)den);
    TRP_LAUNCHED(ctx);
  }
  TRP_TRY(trp_batch_invert_impl(ctx, d->field, den, num, den, n));
  return trp_grand_product_impl(ctx, d->field, den, n, d_last_z, d_z, n, tiles);
}

// z (n_out values, z[0] = 1) of one lookup argument
int trp_lookup_product_impl(trp_domain* d, const void* d_a, const void* d_s, const void* d_ap, const void* d_sp,
                            const uint64_t* consts_host /* beta, gamma */, void* d_z, size_t n_out, void* ws) {
  trp_ctx* ctx = d->ctx;
  const size_t n = (size_t)1 << d->k;
  if (n_out > n) TRP_FAIL(ctx, TRP_E_INVALID, "lookup product of %zu rows exceeds the domain size %zu", n_out, n);
  char* p = (char*)ws;
  void* num = p; p += ws_align(n * 32);
  void* den = p; p += ws_align(n * 32);
  void* tiles = p; p += gp_tiles_bytes(n);
  void* dconsts = p;
  TRP_CUDA(ctx, cudaMemcpyAsync(dconsts, consts_host, 2 * 32, cudaMemcpyHostToDevice, ctx->stream));
  unsigned blocks = (unsigned)((n + 127) / 128);
  {
    ProfScope ps(ctx, PROF_PRODUCTS);
    if (d->field == 0) lookup_terms_kernel<FpParams><<<blocks, 128, 0, ctx->stream>>>((const uint4*)d_a, (const uint4*)d_s, (const uint4*)d_ap, (const uint4*)d_sp, (const uint4*)dconsts, n, (uint4*)num, (uint4*)den);
    else lookup_terms_kernel<FqParams><<<blocks, 128, 0, ctx->stream>>>((const uint4*)d_a, (const uint4*)d_s, (const uint4*)d_ap, (const uint4*)d_sp, (const uint4*)dconsts, n, (uint4*)num, (uint4*)den);
    TRP_LAUNCHED(ctx);
  }
  TRP_TRY(trp_batch_invert_impl(ctx, d->field, den, num, den, n));
  return trp_grand_product_impl(ctx, d->field, den, n, nullptr, d_z, n_out, tiles);
}
