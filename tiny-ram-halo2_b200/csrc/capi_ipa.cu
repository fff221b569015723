// extern "C" surface of SURVEY.md 8(f) row f2 (include/tr_prover.h, section "opening phase"): evaluations, inner products,
// Kate division and the building blocks of the inner-product-argument rounds.  Kernels: ipa.cu, products.cu, msm.cu.
#include "common.cuh"

using namespace ff;

namespace {

struct Locked {
  std::lock_guard<std::mutex> g;
  explicit Locked(trp_ctx* c) : g(c->mu) { cudaSetDevice(c->device); }
};

inline int field_id(const trp_ctx* ctx, int which_field) {
  return which_field == 0 ? scalar_field_of(ctx->curve) : base_field_of(ctx->curve);
}

template <class PR> Fe<PR> fe_from_u64x4(const uint64_t* l) {
  Fe<PR> r;
  for (int i = 0; i < 4; ++i) { r.v[2 * i] = (uint32_t)l[i]; r.v[2 * i + 1] = (uint32_t)(l[i] >> 32); }
  return r;
}
template <class PR> void fe_to_u64x4(const Fe<PR>& a, uint64_t* l) {
  for (int i = 0; i < 4; ++i) l[i] = (uint64_t)a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32);
}
// host-side field helpers on Montgomery limbs
void host_inverse(int field, const uint64_t in[4], uint64_t out[4]) {
  if (field == 0) fe_to_u64x4(fe_inv(fe_from_u64x4<FpParams>(in)), out); else fe_to_u64x4(fe_inv(fe_from_u64x4<FqParams>(in)), out);
}
void host_from_mont(int field, const uint64_t in[4], uint64_t out[4]) {
  if (field == 0) fe_to_u64x4(fe_from_mont(fe_from_u64x4<FpParams>(in)), out); else fe_to_u64x4(fe_from_mont(fe_from_u64x4<FqParams>(in)), out);
}
inline bool is_zero4(const uint64_t* v) { return (v[0] | v[1] | v[2] | v[3]) == 0; }

}  // namespace

extern "C" {

int trp_dev_eval_polynomials(trp_ctx* ctx, int which_field, const uint64_t* d_polys, size_t stride, size_t n, size_t m,
                             const uint64_t x[4], uint64_t* d_out) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (m == 0) return TRP_OK;
  if ((n && !d_polys) || !x || !d_out) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  if (m > 1 && stride < n) TRP_FAIL(ctx, TRP_E_INVALID, "stride %zu is smaller than the polynomial length %zu", stride, n);
  TRP_TRY(trp_ws_reserve(ctx, trp_reduce_ws_bytes(ctx, n, m)));
  return trp_eval_polys_impl(ctx, field_id(ctx, which_field), d_polys, stride, nullptr, n, m, x, d_out, ctx->ws);
}

int trp_dev_eval_polynomials_at(trp_ctx* ctx, int which_field, const uint64_t* const* d_poly_ptrs, size_t n, size_t m, const uint64_t x[4],
                                uint64_t* d_out) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (m == 0) return TRP_OK;
  if (!d_poly_ptrs || !x || !d_out) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  for (size_t j = 0; n && j < m; ++j)
    if (!d_poly_ptrs[j]) TRP_FAIL(ctx, TRP_E_INVALID, "polynomial %zu is NULL", j);
  const size_t rb = trp_reduce_ws_bytes(ctx, n, m), tb = ws_align(m * sizeof(void*));
  TRP_TRY(trp_ws_reserve(ctx, rb + tb));
  void* d_tab = (char*)ctx->ws + rb;            // the pointer table rides in the workspace behind the reduction scratch
  TRP_CUDA(ctx, cudaMemcpyAsync(d_tab, d_poly_ptrs, m * sizeof(void*), cudaMemcpyHostToDevice, ctx->stream));
  int rc = trp_eval_polys_impl(ctx, field_id(ctx, which_field), nullptr, 0, (const void* const*)d_tab, n, m, x, d_out, ctx->ws);
  TRP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // d_poly_ptrs is a pageable host array: the copy must have left it
  return rc;
}

int trp_dev_linear_combination(trp_ctx* ctx, int which_field, const uint64_t* const* d_poly_ptrs, const uint64_t* scalars, size_t n, size_t m,
                               uint64_t* d_out) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (!d_out && n) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  if (m == 0) { if (n) TRP_CUDA(ctx, cudaMemsetAsync(d_out, 0, n * 32, ctx->stream)); return TRP_OK; }
  if (!d_poly_ptrs || !scalars) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  if (m > 0xffffffffu) TRP_FAIL(ctx, TRP_E_INVALID, "too many terms");
  for (size_t j = 0; n && j < m; ++j) {
    if (!d_poly_ptrs[j]) TRP_FAIL(ctx, TRP_E_INVALID, "polynomial %zu is NULL", j);
    if (d_poly_ptrs[j] == d_out) TRP_FAIL(ctx, TRP_E_INVALID, "the output may not alias input %zu", j);
  }
  const size_t tb = ws_align(m * sizeof(void*)), sb = ws_align(m * 32);
  TRP_TRY(trp_ws_reserve(ctx, tb + sb));
  char* d_tab = (char*)ctx->ws; char* d_scal = d_tab + tb;
  TRP_CUDA(ctx, cudaMemcpyAsync(d_tab, d_poly_ptrs, m * sizeof(void*), cudaMemcpyHostToDevice, ctx->stream));
  TRP_CUDA(ctx, cudaMemcpyAsync(d_scal, scalars, m * 32, cudaMemcpyHostToDevice, ctx->stream));
  int rc = trp_lincomb_impl(ctx, field_id(ctx, which_field), (const void* const*)d_tab, d_scal, m, n, d_out);
  TRP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the host arrays are pageable: the copies must have left them
  return rc;
}

int trp_eval_polynomial(trp_ctx* ctx, int which_field, const uint64_t* coeffs, size_t n, const uint64_t x[4], uint64_t out[4]) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if ((n && !coeffs) || !x || !out) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  size_t rb = trp_reduce_ws_bytes(ctx, n, 1), cb = ws_align((n ? n : 1) * 32);
  TRP_TRY(trp_ws_reserve(ctx, rb + cb + 256));
  char* dc = (char*)ctx->ws + rb; char* dout = dc + cb;
  if (n) TRP_CUDA(ctx, cudaMemcpyAsync(dc, coeffs, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  TRP_TRY(trp_eval_polys_impl(ctx, field_id(ctx, which_field), dc, n, nullptr, n, 1, x, dout, ctx->ws));
  TRP_CUDA(ctx, cudaMemcpyAsync(out, dout, 32, cudaMemcpyDeviceToHost, ctx->stream));
  TRP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TRP_OK;
}

int trp_dev_inner_products(trp_ctx* ctx, int which_field, const uint64_t* d_a, size_t a_stride, const uint64_t* d_b, size_t b_stride,
                           size_t n, size_t m, uint64_t* d_out) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (m == 0) return TRP_OK;
  if ((n && (!d_a || !d_b)) || !d_out) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  TRP_TRY(trp_ws_reserve(ctx, trp_reduce_ws_bytes(ctx, n, m)));
  return trp_inner_products_impl(ctx, field_id(ctx, which_field), d_a, a_stride, d_b, b_stride, n, m, d_out, ctx->ws);
}

int trp_compute_inner_product(trp_ctx* ctx, int which_field, const uint64_t* a, const uint64_t* b, size_t n, uint64_t out[4]) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if ((n && (!a || !b)) || !out) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  size_t rb = trp_reduce_ws_bytes(ctx, n, 1), cb = ws_align((n ? n : 1) * 32);
  TRP_TRY(trp_ws_reserve(ctx, rb + 2 * cb + 256));
  char* da = (char*)ctx->ws + rb; char* db = da + cb; char* dout = db + cb;
  if (n) {
    TRP_CUDA(ctx, cudaMemcpyAsync(da, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    TRP_CUDA(ctx, cudaMemcpyAsync(db, b, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  }
  TRP_TRY(trp_inner_products_impl(ctx, field_id(ctx, which_field), da, n, db, n, n, 1, dout, ctx->ws));
  TRP_CUDA(ctx, cudaMemcpyAsync(out, dout, 32, cudaMemcpyDeviceToHost, ctx->stream));
  TRP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TRP_OK;
}

int trp_dev_powers(trp_ctx* ctx, int which_field, const uint64_t x[4], size_t n, uint64_t* d_out) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (n && (!x || !d_out)) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  return trp_powers_impl(ctx, field_id(ctx, which_field), x, n, d_out);
}

int trp_dev_fold(trp_ctx* ctx, int which_field, uint64_t* d_a, size_t half, const uint64_t u[4]) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (half && (!d_a || !u)) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  return trp_fold_impl(ctx, field_id(ctx, which_field), d_a, half, u);
}

int trp_dev_ipa_round_scalars(trp_ctx* ctx, int which_field, const uint64_t* d_p, const uint64_t* d_s, size_t cur, size_t lo, size_t count,
                              size_t col_stride, uint64_t* d_out) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (count == 0) return TRP_OK;
  if (!d_p || !d_s || !d_out) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  if (cur < 2 || (cur & (cur - 1)) || col_stride < count) TRP_FAIL(ctx, TRP_E_INVALID, "cur must be a power of two >= 2 and col_stride >= count");
  unsigned cur_log = 0;
  while (((size_t)1 << cur_log) < cur) ++cur_log;
  return trp_ipa_round_scalars_impl(ctx, field_id(ctx, which_field), d_p, d_s, cur_log, lo, count, col_stride, d_out);
}

int trp_dev_ipa_s_double(trp_ctx* ctx, int which_field, const uint64_t* d_s, size_t m, const uint64_t u[4], uint64_t* d_out) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (m == 0) return TRP_OK;
  if (!d_s || !u || !d_out || d_s == d_out) TRP_FAIL(ctx, TRP_E_INVALID, "NULL or aliased buffer");
  return trp_ipa_s_double_impl(ctx, field_id(ctx, which_field), d_s, m, u, d_out);
}

int trp_dev_kate_division(trp_ctx* ctx, int which_field, const uint64_t* d_coeffs, size_t n, const uint64_t b[4], uint64_t* d_q) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (n <= 1) return TRP_OK;
  if (!d_coeffs || !b || !d_q) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  const int field = field_id(ctx, which_field);
  uint64_t binv[4] = {0, 0, 0, 0};
  const bool zero = is_zero4(b);
  if (!zero) host_inverse(field, b, binv);
  TRP_TRY(trp_ws_reserve(ctx, trp_kate_ws_bytes(n)));
  return trp_kate_division_impl(ctx, field, d_coeffs, n, b, binv, zero, d_q, ctx->ws);
}

int trp_kate_division(trp_ctx* ctx, int which_field, const uint64_t* coeffs, size_t n, const uint64_t b[4], uint64_t* q) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (n <= 1) return TRP_OK;
  if (!coeffs || !b || !q) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  const int field = field_id(ctx, which_field);
  uint64_t binv[4] = {0, 0, 0, 0};
  const bool zero = is_zero4(b);
  if (!zero) host_inverse(field, b, binv);
  size_t kb = trp_kate_ws_bytes(n), cb = ws_align(n * 32);
  TRP_TRY(trp_ws_reserve(ctx, kb + 2 * cb));
  char* dc = (char*)ctx->ws + kb; char* dq = dc + cb;
  TRP_CUDA(ctx, cudaMemcpyAsync(dc, coeffs, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  TRP_TRY(trp_kate_division_impl(ctx, field, dc, n, b, binv, zero, dq, ctx->ws));
  TRP_CUDA(ctx, cudaMemcpyAsync(q, dq, (n - 1) * 32, cudaMemcpyDeviceToHost, ctx->stream));
  TRP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TRP_OK;
}

int trp_dev_generator_collapse(trp_ctx* ctx, uint64_t* d_g, size_t half, const uint64_t u[4]) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (half == 0) return TRP_OK;
  if (!d_g || !u) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  uint64_t canon[4];
  host_from_mont(scalar_field_of(ctx->curve), u, canon);
  TRP_TRY(trp_ws_reserve(ctx, trp_collapse_ws_bytes(half)));
  return trp_generator_collapse_impl(ctx, d_g, half, canon, ctx->ws);
}

int trp_dev_msm_var(trp_ctx* ctx, const uint64_t* d_bases, const uint64_t* d_scalars, size_t n, size_t m, uint64_t* d_out_jacobian) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (m == 0) return TRP_OK;
  if ((n && (!d_bases || !d_scalars)) || !d_out_jacobian) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  if (n > ((size_t)1 << 27)) TRP_FAIL(ctx, TRP_E_INVALID, "too many bases (%zu)", n);
  TRP_TRY(trp_ws_reserve(ctx, trp_msm_var_ws_bytes(n, m)));
  return trp_msm_var_impl(ctx, d_bases, d_scalars, n, m, d_out_jacobian, ctx->ws, ctx->ws_bytes);
}

}  // extern "C"
