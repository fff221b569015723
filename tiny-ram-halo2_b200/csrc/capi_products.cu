// extern "C" surface of SURVEY.md 8(f) row f1 (include/tr_prover.h, section "between the hot kernels"): batch inversion,
// grand products, the permutation / lookup Z columns and permute_expression_pair.  Argument checks and staging only; the
// kernels are in products.cu and lookup.cu.
#include "common.cuh"

#include <vector>

namespace {

struct Locked {
  std::lock_guard<std::mutex> g;
  explicit Locked(trp_ctx* c) : g(c->mu) { cudaSetDevice(c->device); }
};

inline int field_id(const trp_ctx* ctx, int which_field) {
  return which_field == 0 ? scalar_field_of(ctx->curve) : base_field_of(ctx->curve);
}

}  // namespace

extern "C" {

int trp_dev_batch_invert(trp_ctx* ctx, int which_field, const uint64_t* d_a, const uint64_t* d_mul, uint64_t* d_out, size_t n) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (n == 0) return TRP_OK;
  if (!d_a || !d_out) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  return trp_batch_invert_impl(ctx, field_id(ctx, which_field), d_a, d_mul, d_out, n);
}

int trp_batch_invert(trp_ctx* ctx, int which_field, uint64_t* a, size_t n) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (n == 0) return TRP_OK;
  if (!a) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  TRP_TRY(trp_ws_reserve(ctx, ws_align(n * 32)));
  TRP_CUDA(ctx, cudaMemcpyAsync(ctx->ws, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  TRP_TRY(trp_batch_invert_impl(ctx, field_id(ctx, which_field), ctx->ws, nullptr, ctx->ws, n));
  TRP_CUDA(ctx, cudaMemcpyAsync(a, ctx->ws, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  TRP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TRP_OK;
}

int trp_dev_grand_product(trp_ctx* ctx, int which_field, const uint64_t* d_v, size_t n_in, const uint64_t* d_init, uint64_t* d_z,
                          size_t n_out) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (n_out == 0) return TRP_OK;
  if ((n_in && !d_v) || !d_z) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  if (n_out > n_in + 1) TRP_FAIL(ctx, TRP_E_INVALID, "grand product of %zu values yields at most %zu outputs (asked for %zu)", n_in, n_in + 1, n_out);
  TRP_TRY(trp_ws_reserve(ctx, trp_grand_product_ws_bytes(n_out)));
  return trp_grand_product_impl(ctx, field_id(ctx, which_field), d_v, n_in, d_init, d_z, n_out, ctx->ws);
}

int trp_grand_product(trp_ctx* ctx, int which_field, const uint64_t* v, size_t n_in, const uint64_t init[4], uint64_t* z, size_t n_out) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (n_out == 0) return TRP_OK;
  if ((n_in && !v) || !z) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  if (n_out > n_in + 1) TRP_FAIL(ctx, TRP_E_INVALID, "grand product of %zu values yields at most %zu outputs (asked for %zu)", n_in, n_in + 1, n_out);
  size_t vb = ws_align((n_in ? n_in : 1) * 32), zb = ws_align(n_out * 32), tb = trp_grand_product_ws_bytes(n_out);
  TRP_TRY(trp_ws_reserve(ctx, vb + zb + tb + 256));
  char* dv = (char*)ctx->ws; char* dz = dv + vb; char* dt = dz + zb; char* di = dt + tb;
  if (n_in) TRP_CUDA(ctx, cudaMemcpyAsync(dv, v, n_in * 32, cudaMemcpyHostToDevice, ctx->stream));
  if (init) TRP_CUDA(ctx, cudaMemcpyAsync(di, init, 32, cudaMemcpyHostToDevice, ctx->stream));
  TRP_TRY(trp_grand_product_impl(ctx, field_id(ctx, which_field), dv, n_in, init ? di : nullptr, dz, n_out, dt));
  TRP_CUDA(ctx, cudaMemcpyAsync(z, dz, n_out * 32, cudaMemcpyDeviceToHost, ctx->stream));
  TRP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TRP_OK;
}

static int pack_perm_consts(trp_ctx* ctx, size_t m, const uint64_t beta[4], const uint64_t gamma[4], const uint64_t* delta_beta,
                            std::vector<uint64_t>& out) {
  if (!beta || !gamma || !delta_beta) TRP_FAIL(ctx, TRP_E_INVALID, "NULL challenge");
  out.resize((2 + m) * 4);
  for (int i = 0; i < 4; ++i) { out[i] = beta[i]; out[4 + i] = gamma[i]; }
  for (size_t c = 0; c < 4 * m; ++c) out[8 + c] = delta_beta[c];
  return TRP_OK;
}

int trp_dev_permutation_product(trp_domain* d, const uint64_t* const* d_values, const uint64_t* const* d_sigmas, size_t m,
                                const uint64_t beta[4], const uint64_t gamma[4], const uint64_t* delta_beta, const uint64_t* d_last_z,
                                uint64_t* d_z) {
  if (!d) return TRP_E_INVALID;
  trp_ctx* ctx = d->ctx;
  Locked l(ctx);
  if (!d_values || !d_sigmas || !d_z) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  std::vector<uint64_t> consts;
  TRP_TRY(pack_perm_consts(ctx, m, beta, gamma, delta_beta, consts));
  TRP_TRY(trp_ws_reserve(ctx, trp_product_ws_bytes((size_t)1 << d->k)));
  return trp_permutation_product_impl(d, d_values, d_sigmas, m, consts.data(), d_last_z, d_z, ctx->ws);
}

int trp_permutation_product(trp_domain* d, const uint64_t* const* values, const uint64_t* const* sigmas, size_t m,
                            const uint64_t beta[4], const uint64_t gamma[4], const uint64_t* delta_beta, const uint64_t last_z[4],
                            uint64_t* z) {
  if (!d) return TRP_E_INVALID;
  trp_ctx* ctx = d->ctx;
  Locked l(ctx);
  if (!values || !sigmas || !z) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  if (m == 0 || m > 16) TRP_FAIL(ctx, TRP_E_INVALID, "a permutation chunk holds 1..16 columns (got %zu)", m);
  std::vector<uint64_t> consts;
  TRP_TRY(pack_perm_consts(ctx, m, beta, gamma, delta_beta, consts));
  const size_t n = (size_t)1 << d->k, cb = ws_align(n * 32), pw = trp_product_ws_bytes(n);
  TRP_TRY(trp_ws_reserve(ctx, pw + (2 * m + 1) * cb + 256));
  char* base = (char*)ctx->ws + pw;
  std::vector<const uint64_t*> dv(m), ds(m);
  for (size_t c = 0; c < m; ++c) {
    if (!values[c] || !sigmas[c]) TRP_FAIL(ctx, TRP_E_INVALID, "NULL column pointer in permutation chunk");
    dv[c] = (const uint64_t*)(base + (2 * c) * cb); ds[c] = (const uint64_t*)(base + (2 * c + 1) * cb);
    TRP_CUDA(ctx, cudaMemcpyAsync((void*)dv[c], values[c], n * 32, cudaMemcpyHostToDevice, ctx->stream));
    TRP_CUDA(ctx, cudaMemcpyAsync((void*)ds[c], sigmas[c], n * 32, cudaMemcpyHostToDevice, ctx->stream));
  }
  char* dz = base + 2 * m * cb; char* dl = dz + cb;
  if (last_z) TRP_CUDA(ctx, cudaMemcpyAsync(dl, last_z, 32, cudaMemcpyHostToDevice, ctx->stream));
  TRP_TRY(trp_permutation_product_impl(d, dv.data(), ds.data(), m, consts.data(), last_z ? dl : nullptr, dz, ctx->ws));
  TRP_CUDA(ctx, cudaMemcpyAsync(z, dz, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  TRP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TRP_OK;
}

int trp_dev_lookup_product(trp_domain* d, const uint64_t* d_input, const uint64_t* d_table, const uint64_t* d_perm_input,
                           const uint64_t* d_perm_table, const uint64_t beta[4], const uint64_t gamma[4], uint64_t* d_z, size_t n_out) {
  if (!d) return TRP_E_INVALID;
  trp_ctx* ctx = d->ctx;
  Locked l(ctx);
  if (!d_input || !d_table || !d_perm_input || !d_perm_table || !d_z || !beta || !gamma) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  uint64_t consts[8];
  for (int i = 0; i < 4; ++i) { consts[i] = beta[i]; consts[4 + i] = gamma[i]; }
  TRP_TRY(trp_ws_reserve(ctx, trp_product_ws_bytes((size_t)1 << d->k)));
  return trp_lookup_product_impl(d, d_input, d_table, d_perm_input, d_perm_table, consts, d_z, n_out, ctx->ws);
}

int trp_lookup_product(trp_domain* d, const uint64_t* input, const uint64_t* table, const uint64_t* perm_input,
                       const uint64_t* perm_table, const uint64_t beta[4], const uint64_t gamma[4], uint64_t* z, size_t n_out) {
  if (!d) return TRP_E_INVALID;
  trp_ctx* ctx = d->ctx;
  Locked l(ctx);
  if (!input || !table || !perm_input || !perm_table || !z || !beta || !gamma) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  uint64_t consts[8];
  for (int i = 0; i < 4; ++i) { consts[i] = beta[i]; consts[4 + i] = gamma[i]; }
  const size_t n = (size_t)1 << d->k, cb = ws_align(n * 32), pw = trp_product_ws_bytes(n);
  if (n_out > n) TRP_FAIL(ctx, TRP_E_INVALID, "lookup product of %zu rows exceeds the domain size %zu", n_out, n);
  TRP_TRY(trp_ws_reserve(ctx, pw + 5 * cb));
  char* base = (char*)ctx->ws + pw;
  const uint64_t* src[4] = {input, table, perm_input, perm_table};
  for (int c = 0; c < 4; ++c) TRP_CUDA(ctx, cudaMemcpyAsync(base + c * cb, src[c], n * 32, cudaMemcpyHostToDevice, ctx->stream));
  char* dz = base + 4 * cb;
  TRP_TRY(trp_lookup_product_impl(d, base, base + cb, base + 2 * cb, base + 3 * cb, consts, dz, n_out, ctx->ws));
  TRP_CUDA(ctx, cudaMemcpyAsync(z, dz, n_out * 32, cudaMemcpyDeviceToHost, ctx->stream));
  TRP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TRP_OK;
}

int trp_dev_permute_expression_pair(trp_ctx* ctx, const uint64_t* d_input, const uint64_t* d_table, size_t rows,
                                    uint64_t* d_perm_input, uint64_t* d_perm_table, int* all_found) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (rows && (!d_input || !d_table || !d_perm_input || !d_perm_table)) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  TRP_TRY(trp_ws_reserve(ctx, trp_permute_pair_ws_bytes(rows)));
  return trp_permute_pair_impl(ctx, scalar_field_of(ctx->curve), d_input, d_table, rows, d_perm_input, d_perm_table, ctx->ws, all_found);
}

int trp_permute_expression_pair(trp_ctx* ctx, const uint64_t* input, const uint64_t* table, size_t rows, uint64_t* perm_input,
                                uint64_t* perm_table, int* all_found) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (all_found) *all_found = 1;
  if (rows == 0) return TRP_OK;
  if (!input || !table || !perm_input || !perm_table) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  const size_t cb = ws_align(rows * 32), pw = trp_permute_pair_ws_bytes(rows);
  TRP_TRY(trp_ws_reserve(ctx, pw + 4 * cb));
  char* base = (char*)ctx->ws + pw;
  TRP_CUDA(ctx, cudaMemcpyAsync(base, input, rows * 32, cudaMemcpyHostToDevice, ctx->stream));
  TRP_CUDA(ctx, cudaMemcpyAsync(base + cb, table, rows * 32, cudaMemcpyHostToDevice, ctx->stream));
  TRP_TRY(trp_permute_pair_impl(ctx, scalar_field_of(ctx->curve), base, base + cb, rows, base + 2 * cb, base + 3 * cb, ctx->ws, all_found));
  TRP_CUDA(ctx, cudaMemcpyAsync(perm_input, base + 2 * cb, rows * 32, cudaMemcpyDeviceToHost, ctx->stream));
  TRP_CUDA(ctx, cudaMemcpyAsync(perm_table, base + 3 * cb, rows * 32, cudaMemcpyDeviceToHost, ctx->stream));
  TRP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TRP_OK;
}

}  // extern "C"
