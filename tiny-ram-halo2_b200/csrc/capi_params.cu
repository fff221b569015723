// extern "C" surface of SURVEY.md 8(f) row f3 (include/tr_prover.h, section "Params::new"): hash-to-curve, the group FFT and
// Params::new itself.  Kernels: params.cu.
#include "common.cuh"

using namespace ff;

size_t trp_group_fft_ws_bytes(unsigned log_n);
int trp_group_fft_impl(trp_ctx* ctx, void* d_points, unsigned log_n, const uint64_t omega[4], const uint64_t* scale, void* ws);
int trp_random_field_impl(trp_ctx* ctx, int field, const uint8_t key[32], uint64_t first, size_t n, void* d_out);
int trp_hash_to_curve_impl(trp_ctx* ctx, const char* domain_prefix, const uint8_t* d_msgs, size_t msg_len, const uint8_t* msg_prefix,
                           size_t prefix_len, int append_index, uint64_t first_index, size_t n, void* d_out);

namespace {

struct Locked {
  std::lock_guard<std::mutex> g;
  explicit Locked(trp_ctx* c) : g(c->mu) { cudaSetDevice(c->device); }
};

const uint64_t ROOT_FP[4] = {0xbdad6fabd87ea32fULL, 0xea322bf2b7bb7584ULL, 0x362120830561f81aULL, 0x2bce74deac30ebdaULL};
const uint64_t ROOT_FQ[4] = {0xa70e2c1102b6d05fULL, 0x9bb97ea3c106f049ULL, 0x9e5c4dfd492ae26eULL, 0x2de6a9b8746d3f58ULL};

template <class PR> Fe<PR> fe_from_u64x4(const uint64_t* l) {
  Fe<PR> r;
  for (int i = 0; i < 4; ++i) { r.v[2 * i] = (uint32_t)l[i]; r.v[2 * i + 1] = (uint32_t)(l[i] >> 32); }
  return r;
}
template <class PR> void fe_to_u64x4(const Fe<PR>& a, uint64_t* l) {
  for (int i = 0; i < 4; ++i) l[i] = (uint64_t)a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32);
}

// alpha_inv = ROOT_OF_UNITY_INV^(2^(S - k)) and minv = TWO_INV^k of the curve's scalar field, Montgomery limbs
template <class PR>
void ifft_constants(const uint64_t* root_canon, unsigned k, uint64_t alpha_inv[4], uint64_t minv[4]) {
  Fe<PR> w = fe_inv(fe_to_mont(fe_from_u64x4<PR>(root_canon)));
  for (unsigned i = k; i < 32; ++i) w = fe_sqr(w);
  fe_to_u64x4(w, alpha_inv);
  Fe<PR> two = fe_dbl(fe_one<PR>());
  Fe<PR> ti = fe_inv(two), m = fe_one<PR>();
  for (unsigned i = 0; i < k; ++i) m = fe_mul(m, ti);
  fe_to_u64x4(m, minv);
}

// g (device, n points), g_lagrange (device, n points), wu (device, 2 points)
int params_new_device(trp_ctx* ctx, unsigned k, uint64_t* d_g, uint64_t* d_gl, uint64_t* d_wu) {
  const size_t n = (size_t)1 << k;
  const uint8_t zero = 0, one = 1, two = 2;
  TRP_TRY(trp_hash_to_curve_impl(ctx, "Halo2-Parameters", nullptr, 0, &zero, 1, 1, 0, n, d_g));
  TRP_TRY(trp_hash_to_curve_impl(ctx, "Halo2-Parameters", nullptr, 0, &one, 1, 0, 0, 1, d_wu));
  TRP_TRY(trp_hash_to_curve_impl(ctx, "Halo2-Parameters", nullptr, 0, &two, 1, 0, 0, 1, d_wu + 8));
  TRP_CUDA(ctx, cudaMemcpyAsync(d_gl, d_g, n * 64, cudaMemcpyDeviceToDevice, ctx->stream));
  uint64_t alpha_inv[4], minv[4];
  if (scalar_field_of(ctx->curve) == 0) ifft_constants<FpParams>(ROOT_FP, k, alpha_inv, minv);
  else ifft_constants<FqParams>(ROOT_FQ, k, alpha_inv, minv);
  TRP_TRY(trp_ws_reserve(ctx, trp_group_fft_ws_bytes(k)));
  return trp_group_fft_impl(ctx, d_gl, k, alpha_inv, minv, ctx->ws);
}

}  // namespace

extern "C" {

int trp_dev_random_field(trp_ctx* ctx, int which_field, const uint8_t key[32], uint64_t first_counter, size_t n, uint64_t* d_out) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (!key || (n && !d_out)) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  const int field = which_field == 0 ? scalar_field_of(ctx->curve) : base_field_of(ctx->curve);
  return trp_random_field_impl(ctx, field, key, first_counter, n, d_out);
}

int trp_dev_hash_to_curve(trp_ctx* ctx, const char* domain_prefix, const uint8_t* msg_prefix, size_t prefix_len, int append_index,
                          uint64_t first_index, size_t n, uint64_t* d_out) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (!domain_prefix || (prefix_len && !msg_prefix) || (n && !d_out)) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  if (append_index && first_index + n > ((uint64_t)1 << 32)) TRP_FAIL(ctx, TRP_E_INVALID, "the message index is a u32");
  return trp_hash_to_curve_impl(ctx, domain_prefix, nullptr, 0, msg_prefix, prefix_len, append_index, first_index, n, d_out);
}

int trp_hash_to_curve(trp_ctx* ctx, const char* domain_prefix, const uint8_t* messages, size_t msg_len, size_t n, uint64_t* out) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (!domain_prefix || (n && msg_len && !messages) || (n && !out)) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  if (n == 0) return TRP_OK;
  const size_t mb = ws_align(n * msg_len + 1), ob = ws_align(n * 64);
  TRP_TRY(trp_ws_reserve(ctx, mb + ob));
  uint8_t* dm = (uint8_t*)ctx->ws; char* dout = (char*)ctx->ws + mb;
  if (msg_len) TRP_CUDA(ctx, cudaMemcpyAsync(dm, messages, n * msg_len, cudaMemcpyHostToDevice, ctx->stream));
  TRP_TRY(trp_hash_to_curve_impl(ctx, domain_prefix, dm, msg_len, nullptr, 0, 0, 0, n, dout));
  TRP_CUDA(ctx, cudaMemcpyAsync(out, dout, n * 64, cudaMemcpyDeviceToHost, ctx->stream));
  TRP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TRP_OK;
}

int trp_dev_group_fft(trp_ctx* ctx, uint64_t* d_points, unsigned log_n, const uint64_t omega[4], const uint64_t* scale) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (!d_points || !omega) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  if (log_n > 28) TRP_FAIL(ctx, TRP_E_INVALID, "log_n %u is out of range", log_n);
  TRP_TRY(trp_ws_reserve(ctx, trp_group_fft_ws_bytes(log_n)));
  return trp_group_fft_impl(ctx, d_points, log_n, omega, scale, ctx->ws);
}

int trp_group_fft(trp_ctx* ctx, uint64_t* points, unsigned log_n, const uint64_t omega[4], const uint64_t* scale) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (!points || !omega) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  if (log_n > 28) TRP_FAIL(ctx, TRP_E_INVALID, "log_n %u is out of range", log_n);
  const size_t n = (size_t)1 << log_n, fb = trp_group_fft_ws_bytes(log_n);
  TRP_TRY(trp_ws_reserve(ctx, fb + ws_align(n * 64)));
  char* dp = (char*)ctx->ws + fb;
  TRP_CUDA(ctx, cudaMemcpyAsync(dp, points, n * 64, cudaMemcpyHostToDevice, ctx->stream));
  TRP_TRY(trp_group_fft_impl(ctx, dp, log_n, omega, scale, ctx->ws));
  TRP_CUDA(ctx, cudaMemcpyAsync(points, dp, n * 64, cudaMemcpyDeviceToHost, ctx->stream));
  TRP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TRP_OK;
}

int trp_dev_params_new(trp_ctx* ctx, unsigned k, uint64_t* d_g, uint64_t* d_g_lagrange, uint64_t* d_wu) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (!d_g || !d_g_lagrange || !d_wu) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  if (k >= 29) TRP_FAIL(ctx, TRP_E_INVALID, "k = %u is out of range (halo2 asserts k < 32; device memory bounds it earlier)", k);
  return params_new_device(ctx, k, d_g, d_g_lagrange, d_wu);
}

int trp_params_new(trp_ctx* ctx, unsigned k, uint64_t* g, uint64_t* g_lagrange, uint64_t w[8], uint64_t u[8]) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (!g || !g_lagrange || !w || !u) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  if (k >= 29) TRP_FAIL(ctx, TRP_E_INVALID, "k = %u is out of range (halo2 asserts k < 32; device memory bounds it earlier)", k);
  const size_t n = (size_t)1 << k;
  // the group FFT's own scratch sits in ctx->ws; the three outputs get a separate allocation so that reserving cannot move them
  char* buf = nullptr;
  TRP_CUDA(ctx, cudaMalloc(&buf, 2 * n * 64 + 128));
  uint64_t* dg = (uint64_t*)buf; uint64_t* dgl = dg + n * 8; uint64_t* dwu = dgl + n * 8;
  int rc = params_new_device(ctx, k, dg, dgl, dwu);
  cudaError_t e = cudaSuccess;
  if (rc == TRP_OK) {
    e = cudaMemcpyAsync(g, dg, n * 64, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(g_lagrange, dgl, n * 64, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(w, dwu, 64, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(u, dwu + 8, 64, cudaMemcpyDeviceToHost, ctx->stream);
  }
  cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
  cudaFree(buf);
  if (rc != TRP_OK) return rc;
  TRP_CUDA(ctx, e);
  TRP_CUDA(ctx, e2);
  return TRP_OK;
}

}  // extern "C"
