// extern "C" surface declared in include/tr_prover.h: argument checking, host<->device staging, domain constants.
#include "common.cuh"

#include <cstring>
#include <new>

using namespace ff;

// ---- host-side field helpers (ff.cuh compiles for the host too) ---------------------------------------------
namespace {

template <class PR> Fe<PR> fe_from_u64x4(const uint64_t* l) {
  Fe<PR> r;
  for (int i = 0; i < 4; ++i) { r.v[2 * i] = (uint32_t)l[i]; r.v[2 * i + 1] = (uint32_t)(l[i] >> 32); }
  return r;
}
template <class PR> void fe_to_u64x4(const Fe<PR>& a, uint64_t* l) {
  for (int i = 0; i < 4; ++i) l[i] = (uint64_t)a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32);
}

// canonical constants of pasta_curves::fields: ROOT_OF_UNITY (order 2^32) and ZETA
const uint64_t ROOT_FP[4] = {0xbdad6fabd87ea32fULL, 0xea322bf2b7bb7584ULL, 0x362120830561f81aULL, 0x2bce74deac30ebdaULL};
const uint64_t ZETA_FP[4] = {0x1dad5ebdfdfe4ab9ULL, 0x1d1f8bd237ad3149ULL, 0x2caad5dc57aab1b0ULL, 0x12ccca834acdba71ULL};
const uint64_t ROOT_FQ[4] = {0xa70e2c1102b6d05fULL, 0x9bb97ea3c106f049ULL, 0x9e5c4dfd492ae26eULL, 0x2de6a9b8746d3f58ULL};
const uint64_t ZETA_FQ[4] = {0x2aa9d2e050aa0e4fULL, 0x0fed467d47c033afULL, 0x511db4d81cf70f5aULL, 0x06819a58283e528eULL};

struct Locked {
  std::lock_guard<std::mutex> g;
  explicit Locked(trp_ctx* c) : g(c->mu) { cudaSetDevice(c->device); }
};

}  // namespace

namespace {

template <class PR>
int domain_build(trp_ctx* ctx, trp_domain* d, const uint64_t* root_canon, const uint64_t* zeta_canon) {
  const unsigned k = d->k, j = d->j;
  const uint64_t n = 1ULL << k;
  unsigned ext_k = k;
  while ((1ULL << ext_k) < n * (uint64_t)(j - 1)) ++ext_k;
  if (ext_k > 30) TRP_FAIL(ctx, TRP_E_INVALID, "extended_k = %u is out of range", ext_k);
  d->ext_k = ext_k;
  Fe<PR> ext_omega = fe_to_mont(fe_from_u64x4<PR>(root_canon));
  for (unsigned i = ext_k; i < 32; ++i) ext_omega = fe_sqr(ext_omega);
  Fe<PR> omega = ext_omega;
  for (unsigned i = k; i < ext_k; ++i) omega = fe_sqr(omega);
  Fe<PR> zeta = fe_to_mont(fe_from_u64x4<PR>(zeta_canon));
  Fe<PR> zeta2 = fe_sqr(zeta);
  fe_to_u64x4(omega, d->omega); fe_to_u64x4(fe_inv(omega), d->omega_inv);
  fe_to_u64x4(ext_omega, d->ext_omega); fe_to_u64x4(fe_inv(ext_omega), d->ext_omega_inv);
  fe_to_u64x4(zeta, d->g_coset); fe_to_u64x4(zeta2, d->g_coset_inv);
  // t_evaluations: (zeta * ext_omega^i)^n - 1, i < 2^(ext_k-k), inverted
  uint32_t e[2] = {(uint32_t)n, (uint32_t)(n >> 32)};
  Fe<PR> orig = fe_pow(zeta, e, 2), step = fe_pow(ext_omega, e, 2), cur = orig;
  std::vector<Fe<PR>> tabs;
  Fe<PR> one = fe_one<PR>();
  tabs.push_back(one); tabs.push_back(zeta); tabs.push_back(zeta2);                // zeta_in  [0..3)
  Fe<PR> two_k = fe_zero<PR>(); two_k.v[k >> 5] = 1u << (k & 31);
  Fe<PR> ninv = fe_inv(fe_to_mont(two_k));
  tabs.push_back(ninv);                                                           // l2c_post [3]
  Fe<PR> two_ek = fe_zero<PR>(); two_ek.v[ext_k >> 5] = 1u << (ext_k & 31);
  Fe<PR> eninv = fe_inv(fe_to_mont(two_ek));
  tabs.push_back(eninv); tabs.push_back(fe_mul(eninv, zeta2)); tabs.push_back(fe_mul(eninv, zeta));   // e2c_post [4..7)
  d->t_period = 1u << (ext_k - k);
  { Fe<PR> g = zeta; for (unsigned i = 0; i < d->t_period; ++i) { fe_to_u64x4(g, d->coset_gen[i]); g = fe_mul(g, ext_omega); } }
  for (unsigned i = 0; i < d->t_period; ++i) { tabs.push_back(fe_inv(fe_sub(cur, one))); cur = fe_mul(cur, step); }   // tinv [7..)
  TRP_CUDA(ctx, cudaMalloc(&d->d_tabs, tabs.size() * 32));
  TRP_CUDA(ctx, cudaMemcpy(d->d_tabs, tabs.data(), tabs.size() * 32, cudaMemcpyHostToDevice));
  char* base = (char*)d->d_tabs;
  d->d_zeta_in = base; d->d_l2c_post = base + 3 * 32; d->d_e2c_post = base + 4 * 32; d->d_tinv = base + 7 * 32;
  return TRP_OK;
}

// columns per chunk so that scratch stays within `budget` bytes
size_t chunk_cols(size_t batch, size_t bytes_per_col, size_t budget) {
  size_t c = budget / (bytes_per_col ? bytes_per_col : 1);
  if (c < 1) c = 1;
  if (c > batch) c = batch;
  if (c > 65535) c = 65535;
  return c;
}
constexpr size_t SCRATCH_BUDGET = (size_t)4 << 30;

int dev_ntt_inplace(trp_ctx* ctx, int field, uint64_t* d_a, size_t batch, unsigned log_n, const uint64_t omega[4],
                    const void* d_post, unsigned post_period) {
  const size_t N = (size_t)1 << log_n;
  const bool need_tmp = trp_ntt_passes(log_n) > 1;
  size_t cols = need_tmp ? chunk_cols(batch, N * 32, SCRATCH_BUDGET) : (batch > 65535 ? 65535 : batch);
  if (need_tmp) TRP_TRY(trp_ws_reserve(ctx, cols * N * 32));
  for (size_t b0 = 0; b0 < batch; b0 += cols) {
    size_t nb = batch - b0 < cols ? batch - b0 : cols;
    uint64_t* a = d_a + 4 * b0 * N;
    TRP_TRY(trp_ntt_impl(ctx, field, a, a, nb, log_n, omega, N, N, (unsigned)N, nullptr, 0, d_post, post_period,
                         (unsigned)N, need_tmp ? ctx->ws : nullptr));
  }
  return TRP_OK;
}

}  // namespace

// host-pointer helper: stage columns through the arena, run `fn` on the device copy, copy back
template <class F>
static int staged_columns(trp_ctx* ctx, uint64_t* host, size_t batch, size_t elems_per_col, size_t extra_per_col, F fn) {
  size_t col_bytes = elems_per_col * 32;
  size_t cols = chunk_cols(batch, col_bytes + extra_per_col, SCRATCH_BUDGET);
  for (size_t b0 = 0; b0 < batch; b0 += cols) {
    size_t nb = batch - b0 < cols ? batch - b0 : cols;
    TRP_TRY(trp_ws_reserve(ctx, nb * (col_bytes + extra_per_col)));
    char* d = (char*)ctx->ws;
    TRP_CUDA(ctx, cudaMemcpyAsync(d, host + 4 * b0 * elems_per_col, nb * col_bytes, cudaMemcpyHostToDevice, ctx->stream));
    TRP_TRY(fn((uint64_t*)d, nb, d + nb * col_bytes));
    TRP_CUDA(ctx, cudaMemcpyAsync(host + 4 * b0 * elems_per_col, d, nb * col_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    TRP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return TRP_OK;
}


extern "C" {

const char* trp_version(void) { return "tinyram-prover-b200 0.1 (sm_100a)"; }

int trp_ctx_create(trp_ctx** out, int device, int curve) {
  if (!out) return TRP_E_INVALID;
  *out = nullptr;
  if (curve != TRP_CURVE_PALLAS && curve != TRP_CURVE_VESTA) return TRP_E_INVALID;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { cudaGetLastError(); return TRP_E_NODEVICE; }
  if (device < 0 || device >= count) return TRP_E_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return TRP_E_CUDA; }
  trp_ctx* c = new (std::nothrow) trp_ctx();
  if (!c) return TRP_E_OOM;
  c->device = device; c->curve = curve;
  if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); delete c; return TRP_E_CUDA; }
  c->stream = c->own_stream;
  cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
  *out = c;
  return TRP_OK;
}

void trp_ctx_destroy(trp_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  trp_prof_collect(ctx);
  for (auto e : ctx->prof_pool) cudaEventDestroy(e);
  for (auto& t : ctx->twiddles) cudaFree(t.d_tab);
  if (ctx->ws) cudaFree(ctx->ws);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  for (auto& e : ctx->copy_ev) if (e) cudaEventDestroy(e);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

const char* trp_last_error(const trp_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int trp_ctx_set_stream(trp_ctx* ctx, void* cuda_stream) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  TRP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return TRP_OK;
}

int trp_ctx_sync(trp_ctx* ctx) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  TRP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TRP_OK;
}

uint64_t trp_ctx_launch_count(const trp_ctx* ctx) { return ctx ? ctx->launches : 0; }

int trp_prof_enable(trp_ctx* ctx, int on) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  trp_prof_collect(ctx);
  ctx->prof_on = on != 0;
  return TRP_OK;
}
int trp_prof_reset(trp_ctx* ctx) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  trp_prof_collect(ctx);
  for (int i = 0; i < PROF_NPHASES; ++i) { ctx->prof_ms[i] = 0; ctx->prof_count[i] = 0; }
  return TRP_OK;
}
int trp_prof_get(trp_ctx* ctx, int phase, double* total_ms, uint64_t* count) {
  if (!ctx || phase < 0 || phase >= PROF_NPHASES) return TRP_E_INVALID;
  Locked l(ctx);
  trp_prof_collect(ctx);
  if (total_ms) *total_ms = ctx->prof_ms[phase];
  if (count) *count = ctx->prof_count[phase];
  return TRP_OK;
}
int trp_prof_get_work(trp_ctx* ctx, int phase, double* work) {
  if (!ctx || phase < 0 || phase >= PROF_NPHASES || !work) return TRP_E_INVALID;
  Locked l(ctx);
  trp_prof_collect(ctx);
  *work = ctx->prof_work[phase];
  return TRP_OK;
}

// ---- MSM ---------------------------------------------------------------------------------------------------
int trp_bases_load(trp_ctx* ctx, const uint64_t* affine_xy, size_t n, trp_bases** out) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (n && !affine_xy) TRP_FAIL(ctx, TRP_E_INVALID, "affine_xy is NULL");
  return trp_bases_create(ctx, affine_xy, false, n, 0, out);
}
int trp_dev_bases_load(trp_ctx* ctx, const uint64_t* d_affine_xy, size_t n, trp_bases** out) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (n && !d_affine_xy) TRP_FAIL(ctx, TRP_E_INVALID, "d_affine_xy is NULL");
  return trp_bases_create(ctx, d_affine_xy, true, n, 0, out);
}
// flags: bit 0 = keep per-window bucket sets (no precomputed table), bit 1 = force the precomputed table, bits 8..15 = window
// width c (0 = the library's choice for n)
int trp_dev_bases_load_ex(trp_ctx* ctx, const uint64_t* d_affine_xy, size_t n, int flags, trp_bases** out) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (n && !d_affine_xy) TRP_FAIL(ctx, TRP_E_INVALID, "d_affine_xy is NULL");
  return trp_bases_create(ctx, d_affine_xy, true, n, flags, out);
}
int trp_bases_load_ex(trp_ctx* ctx, const uint64_t* affine_xy, size_t n, int flags, trp_bases** out) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (n && !affine_xy) TRP_FAIL(ctx, TRP_E_INVALID, "affine_xy is NULL");
  return trp_bases_create(ctx, affine_xy, false, n, flags, out);
}
size_t trp_bases_len(const trp_bases* b) { return b ? b->n : 0; }
void trp_bases_free(trp_bases* b) {
  if (!b) return;
  Locked l(b->ctx);
  cudaStreamSynchronize(b->ctx->stream);
  trp_bases_destroy(b);
}
// window bits, window count and whether the 2^(c*w) multiples were precomputed (reported by bench.py)
int trp_bases_describe(const trp_bases* b, unsigned out[3]) {
  if (!b || !out) return TRP_E_INVALID;
  return trp_bases_info(b, &out[0], &out[1], &out[2]);
}

int trp_dev_msm_batch(trp_ctx* ctx, const trp_bases* bases, const uint64_t* d_scalars, size_t n, size_t m,
                      uint64_t* d_out_jacobian) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (!bases || bases->ctx != ctx) TRP_FAIL(ctx, TRP_E_INVALID, "bases handle does not belong to this context");
  if (m == 0) return TRP_OK;
  if ((n && !d_scalars) || !d_out_jacobian) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  size_t need = trp_msm_ws_bytes(bases, n, m);
  TRP_TRY(trp_ws_reserve(ctx, need));
  return trp_msm_impl(ctx, bases, d_scalars, n, m, d_out_jacobian, ctx->ws, ctx->ws_bytes);
}

int trp_msm_batch(trp_ctx* ctx, const trp_bases* bases, const uint64_t* scalars, size_t n, size_t m, uint64_t* out_jacobian) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (!bases || bases->ctx != ctx) TRP_FAIL(ctx, TRP_E_INVALID, "bases handle does not belong to this context");
  if (m == 0) return TRP_OK;
  if ((n && !scalars) || !out_jacobian) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  if (n > bases->n) TRP_FAIL(ctx, TRP_E_INVALID, "MSM length %zu exceeds the %zu loaded bases", n, bases->n);
  // stage up to `cols` scalar columns at a time
  size_t col_bytes = ws_align(n * 32 + 32);
  size_t cols = chunk_cols(m, col_bytes, SCRATCH_BUDGET);
  size_t msm_ws = trp_msm_ws_bytes(bases, n, cols);
  size_t out_bytes = ws_align(m * 96);
  TRP_TRY(trp_ws_reserve(ctx, cols * col_bytes + out_bytes + msm_ws));
  char* ws = (char*)ctx->ws;
  char* d_sc = ws; char* d_out = ws + cols * col_bytes; char* d_msm = d_out + out_bytes;
  size_t msm_cap = ctx->ws_bytes - (size_t)(d_msm - ws);
  if (!ctx->copy_stream) {
    TRP_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (auto& e : ctx->copy_ev) TRP_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  for (size_t k0 = 0; k0 < m; k0 += cols) {
    size_t nk = m - k0 < cols ? m - k0 : cols;
    // upload on the copy stream in two parts (1 column, then the rest): the first column's kernels start as soon as it has
    // landed and hide the upload of the others (PCIe moves a column ~5x faster than the GPU consumes it)
    size_t part[2] = {nk >= 3 ? 1 : nk, nk >= 3 ? nk - 1 : 0};
    TRP_CUDA(ctx, cudaEventRecord(ctx->copy_ev[2], ctx->stream));           // staging area is free once earlier work is done
    TRP_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev[2], 0));
    size_t c0 = 0;
    for (int p = 0; p < 2 && part[p]; ++p) {
      if (n) TRP_CUDA(ctx, cudaMemcpyAsync(d_sc + c0 * n * 32, scalars + 4 * (k0 + c0) * n, part[p] * n * 32, cudaMemcpyHostToDevice, ctx->copy_stream));
      TRP_CUDA(ctx, cudaEventRecord(ctx->copy_ev[p], ctx->copy_stream));
      c0 += part[p];
    }
    c0 = 0;
    for (int p = 0; p < 2 && part[p]; ++p) {
      TRP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[p], 0));
      TRP_TRY(trp_msm_impl(ctx, bases, d_sc + c0 * n * 32, n, part[p], d_out + 96 * (k0 + c0), d_msm, msm_cap));
      c0 += part[p];
    }
  }
  TRP_CUDA(ctx, cudaMemcpyAsync(out_jacobian, d_out, m * 96, cudaMemcpyDeviceToHost, ctx->stream));
  TRP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TRP_OK;
}

int trp_msm(trp_ctx* ctx, const trp_bases* bases, const uint64_t* scalars, size_t n, uint64_t out_jacobian[12]) {
  return trp_msm_batch(ctx, bases, scalars, n, 1, out_jacobian);
}

// sum of g group elements given as Jacobian points (normalised result): combines the partial sums of an MSM whose point
// range was split across GPUs (mirrors how best_multiexp adds its per-thread partial results)
int trp_points_sum(trp_ctx* ctx, const uint64_t* jacobian, size_t g, uint64_t out_jacobian[12]) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if ((g && !jacobian) || !out_jacobian) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  TRP_TRY(trp_ws_reserve(ctx, ws_align(g * 96) + 256));
  char* d_in = (char*)ctx->ws; char* d_out = d_in + ws_align(g * 96);
  if (g) TRP_CUDA(ctx, cudaMemcpyAsync(d_in, jacobian, g * 96, cudaMemcpyHostToDevice, ctx->stream));
  TRP_TRY(trp_points_sum_impl(ctx, d_in, g, d_out));
  TRP_CUDA(ctx, cudaMemcpyAsync(out_jacobian, d_out, 96, cudaMemcpyDeviceToHost, ctx->stream));
  TRP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TRP_OK;
}
int trp_dev_points_sum(trp_ctx* ctx, const uint64_t* d_jacobian, size_t g, uint64_t* d_out_jacobian) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if ((g && !d_jacobian) || !d_out_jacobian) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  return trp_points_sum_impl(ctx, d_jacobian, g, d_out_jacobian);
}

// synthetic bases: out[i] = P0 + i*D (affine), written to DEVICE memory
int trp_dev_points_progression(trp_ctx* ctx, const uint64_t p0[8], const uint64_t d[8], size_t n, uint64_t* d_out) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (!p0 || !d || (n && !d_out)) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  return trp_points_progression_impl(ctx, p0, d, n, d_out);
}

// inclusive prefix sums of n affine points on the device: d_out[j] = d_in[0] + ... + d_in[j] (affine, identity = all zero).
// The base set of the summation-by-parts commitment: sum_i z_i G_i = sum_j (z_j - z_{j+1}) Q_j with Q = prefix sums of G.
int trp_dev_points_prefix_sum(trp_ctx* ctx, const uint64_t* d_in, size_t n, uint64_t* d_out) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (n && (!d_in || !d_out)) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  TRP_TRY(trp_ws_reserve(ctx, trp_points_prefix_ws_bytes(n)));
  return trp_points_prefix_sum_impl(ctx, d_in, n, d_out, ctx->ws);
}

// ---- NTT ---------------------------------------------------------------------------------------------------
int trp_dev_ntt(trp_ctx* ctx, uint64_t* d_a, size_t batch, unsigned log_n, const uint64_t omega[4]) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (!omega || (batch && !d_a)) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  if (log_n > 30) TRP_FAIL(ctx, TRP_E_INVALID, "log_n = %u is out of range (max 30)", log_n);
  return dev_ntt_inplace(ctx, scalar_field_of(ctx->curve), d_a, batch, log_n, omega, nullptr, 0);
}

int trp_ntt(trp_ctx* ctx, uint64_t* a, size_t batch, unsigned log_n, const uint64_t omega[4]) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (!omega || (batch && !a)) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  if (log_n > 30) TRP_FAIL(ctx, TRP_E_INVALID, "log_n = %u is out of range (max 30)", log_n);
  const size_t N = (size_t)1 << log_n;
  const int field = scalar_field_of(ctx->curve);
  const bool need_tmp = trp_ntt_passes(log_n) > 1;
  return staged_columns(ctx, a, batch, N, need_tmp ? N * 32 : 0, [&](uint64_t* d, size_t nb, char* tmp) {
    return trp_ntt_impl(ctx, field, d, d, nb, log_n, omega, N, N, (unsigned)N, nullptr, 0, nullptr, 0, (unsigned)N,
                        need_tmp ? tmp : nullptr);
  });
}

// ---- EvaluationDomain ----------------------------------------------------------------------------------------
int trp_domain_create(trp_ctx* ctx, unsigned k, unsigned j, trp_domain** out) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (!out) TRP_FAIL(ctx, TRP_E_INVALID, "out is NULL");
  *out = nullptr;
  if (k > 27 || j < 2 || j > 64) TRP_FAIL(ctx, TRP_E_INVALID, "EvaluationDomain::new(j = %u, k = %u) is out of range", j, k);
  trp_domain* d = new (std::nothrow) trp_domain();
  if (!d) TRP_FAIL(ctx, TRP_E_OOM, "out of host memory");
  d->ctx = ctx; d->k = k; d->j = j; d->d_tabs = nullptr;
  d->field = scalar_field_of(ctx->curve);
  int rc = d->field == 0 ? domain_build<FpParams>(ctx, d, ROOT_FP, ZETA_FP) : domain_build<FqParams>(ctx, d, ROOT_FQ, ZETA_FQ);
  if (rc != TRP_OK) { if (d->d_tabs) cudaFree(d->d_tabs); delete d; return rc; }
  *out = d;
  return TRP_OK;
}
void trp_domain_free(trp_domain* d) {
  if (!d) return;
  Locked l(d->ctx);
  cudaStreamSynchronize(d->ctx->stream);
  if (d->d_tabs) cudaFree(d->d_tabs);
  delete d;
}
unsigned trp_domain_extended_k(const trp_domain* d) { return d ? d->ext_k : 0; }
int trp_domain_constants(const trp_domain* d, uint64_t out[16]) {
  if (!d || !out) return TRP_E_INVALID;
  memcpy(out, d->omega, 32); memcpy(out + 4, d->ext_omega, 32); memcpy(out + 8, d->g_coset, 32); memcpy(out + 12, d->g_coset_inv, 32);
  return TRP_OK;
}

int trp_dev_lagrange_to_coeff(trp_domain* d, uint64_t* d_cols, size_t batch) {
  if (!d) return TRP_E_INVALID;
  trp_ctx* ctx = d->ctx;
  Locked l(ctx);
  if (batch && !d_cols) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  return dev_ntt_inplace(ctx, d->field, d_cols, batch, d->k, d->omega_inv, d->d_l2c_post, 1);
}
int trp_lagrange_to_coeff(trp_domain* d, uint64_t* cols, size_t batch) {
  if (!d) return TRP_E_INVALID;
  trp_ctx* ctx = d->ctx;
  Locked l(ctx);
  if (batch && !cols) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  const size_t N = (size_t)1 << d->k;
  const bool need_tmp = trp_ntt_passes(d->k) > 1;
  return staged_columns(ctx, cols, batch, N, need_tmp ? N * 32 : 0, [&](uint64_t* dc, size_t nb, char* tmp) {
    return trp_ntt_impl(ctx, d->field, dc, dc, nb, d->k, d->omega_inv, N, N, (unsigned)N, nullptr, 0, d->d_l2c_post, 1,
                        (unsigned)N, need_tmp ? tmp : nullptr);
  });
}
int trp_coeff_to_lagrange(trp_domain* d, uint64_t* cols, size_t batch) {
  if (!d) return TRP_E_INVALID;
  return trp_ntt(d->ctx, cols, batch, d->k, d->omega);
}

int trp_dev_coeff_to_extended(trp_domain* d, const uint64_t* d_coeff, uint64_t* d_ext, size_t batch) {
  if (!d) return TRP_E_INVALID;
  trp_ctx* ctx = d->ctx;
  Locked l(ctx);
  if (batch && (!d_coeff || !d_ext)) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  const size_t n = (size_t)1 << d->k, EN = (size_t)1 << d->ext_k;
  for (size_t b0 = 0; b0 < batch; b0 += 65535) {
    size_t nb = batch - b0 < 65535 ? batch - b0 : 65535;
    TRP_TRY(trp_ntt_impl(ctx, d->field, d_coeff + 4 * b0 * n, d_ext + 4 * b0 * EN, nb, d->ext_k, d->ext_omega, n, EN, (unsigned)n,
                         d->d_zeta_in, 3, nullptr, 0, (unsigned)EN, nullptr));
  }
  return TRP_OK;
}
int trp_coeff_to_extended(trp_domain* d, const uint64_t* coeff, uint64_t* ext, size_t batch) {
  if (!d) return TRP_E_INVALID;
  trp_ctx* ctx = d->ctx;
  Locked l(ctx);
  if (batch && (!coeff || !ext)) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  const size_t n = (size_t)1 << d->k, EN = (size_t)1 << d->ext_k;
  size_t cols = chunk_cols(batch, (n + EN) * 32, SCRATCH_BUDGET);
  for (size_t b0 = 0; b0 < batch; b0 += cols) {
    size_t nb = batch - b0 < cols ? batch - b0 : cols;
    TRP_TRY(trp_ws_reserve(ctx, nb * (n + EN) * 32));
    char* d_in = (char*)ctx->ws; char* d_out = d_in + nb * n * 32;
    TRP_CUDA(ctx, cudaMemcpyAsync(d_in, coeff + 4 * b0 * n, nb * n * 32, cudaMemcpyHostToDevice, ctx->stream));
    TRP_TRY(trp_ntt_impl(ctx, d->field, d_in, d_out, nb, d->ext_k, d->ext_omega, n, EN, (unsigned)n, d->d_zeta_in, 3, nullptr, 0,
                         (unsigned)EN, nullptr));
    TRP_CUDA(ctx, cudaMemcpyAsync(ext + 4 * b0 * EN, d_out, nb * EN * 32, cudaMemcpyDeviceToHost, ctx->stream));
    TRP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return TRP_OK;
}

static int dev_extended_to_coeff_locked(trp_domain* d, uint64_t* d_ext, uint64_t* d_out, int divide, void* tmp) {
  trp_ctx* ctx = d->ctx;
  const size_t n = (size_t)1 << d->k, EN = (size_t)1 << d->ext_k;
  const size_t n_out = n * (d->j - 1);
  // single-pass transforms go ext -> out directly; multi-pass ones use tmp as the intermediate buffer
  return trp_ntt_impl(ctx, d->field, d_ext, d_out, 1, d->ext_k, d->ext_omega_inv, EN, n_out, (unsigned)EN,
                      divide ? d->d_tinv : nullptr, d->t_period, d->d_e2c_post, 3, (unsigned)n_out, tmp);
}
int trp_dev_extended_to_coeff(trp_domain* d, uint64_t* d_ext, uint64_t* d_out_coeff, int divide_by_vanishing) {
  if (!d) return TRP_E_INVALID;
  trp_ctx* ctx = d->ctx;
  Locked l(ctx);
  if (!d_ext || !d_out_coeff) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  const size_t EN = (size_t)1 << d->ext_k;
  void* tmp = nullptr;
  if (trp_ntt_passes(d->ext_k) > 1) { TRP_TRY(trp_ws_reserve(ctx, EN * 32)); tmp = ctx->ws; }
  return dev_extended_to_coeff_locked(d, d_ext, d_out_coeff, divide_by_vanishing, tmp);
}
int trp_extended_to_coeff(trp_domain* d, uint64_t* ext, uint64_t* out_coeff, int divide_by_vanishing) {
  if (!d) return TRP_E_INVALID;
  trp_ctx* ctx = d->ctx;
  Locked l(ctx);
  if (!ext || !out_coeff) TRP_FAIL(ctx, TRP_E_INVALID, "NULL buffer");
  const size_t n = (size_t)1 << d->k, EN = (size_t)1 << d->ext_k, n_out = n * (d->j - 1);
  TRP_TRY(trp_ws_reserve(ctx, (2 * EN + n_out) * 32));
  char* d_ext = (char*)ctx->ws; char* d_out = d_ext + EN * 32; char* tmp = d_out + ws_align(n_out * 32);
  TRP_CUDA(ctx, cudaMemcpyAsync(d_ext, ext, EN * 32, cudaMemcpyHostToDevice, ctx->stream));
  TRP_TRY(dev_extended_to_coeff_locked(d, (uint64_t*)d_ext, (uint64_t*)d_out, divide_by_vanishing, trp_ntt_passes(d->ext_k) > 1 ? tmp : nullptr));
  TRP_CUDA(ctx, cudaMemcpyAsync(out_coeff, d_out, n_out * 32, cudaMemcpyDeviceToHost, ctx->stream));
  TRP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TRP_OK;
}

// ---- glue --------------------------------------------------------------------------------------------------
int trp_field_op(trp_ctx* ctx, int which_field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (n == 0) return TRP_OK;
  if (!a || !out || op < 0 || op > 4) TRP_FAIL(ctx, TRP_E_INVALID, "bad argument");
  if ((op <= 2) && !b) TRP_FAIL(ctx, TRP_E_INVALID, "binary op needs b");
  int field = which_field == 0 ? scalar_field_of(ctx->curve) : base_field_of(ctx->curve);
  TRP_TRY(trp_ws_reserve(ctx, 3 * ws_align(n * 32)));
  char* da = (char*)ctx->ws; char* db = da + ws_align(n * 32); char* dout = db + ws_align(n * 32);
  TRP_CUDA(ctx, cudaMemcpyAsync(da, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  if (b) TRP_CUDA(ctx, cudaMemcpyAsync(db, b, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  TRP_TRY(trp_field_op_impl(ctx, field, op, da, b ? db : nullptr, dout, n));
  TRP_CUDA(ctx, cudaMemcpyAsync(out, dout, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  TRP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TRP_OK;
}

int trp_dev_field_op(trp_ctx* ctx, int which_field, int op, const uint64_t* d_a, const uint64_t* d_b, uint64_t* d_out, size_t n) {
  if (!ctx) return TRP_E_INVALID;
  Locked l(ctx);
  if (n == 0) return TRP_OK;
  if (!d_a || !d_out || op < 0 || (op & 15) > 4 || op > 20) TRP_FAIL(ctx, TRP_E_INVALID, "bad argument");
  if (((op & 15) <= 2) && !d_b) TRP_FAIL(ctx, TRP_E_INVALID, "binary op needs b");
  int field = which_field == 0 ? scalar_field_of(ctx->curve) : base_field_of(ctx->curve);
  return trp_field_op_impl(ctx, field, op, d_a, d_b, d_out, n);
}

int trp_microbench(trp_ctx* ctx, int kind, int iters, double* out_gops) {
  if (!ctx || !out_gops) return TRP_E_INVALID;
  Locked l(ctx);
  return trp_microbench_impl(ctx, kind, iters, out_gops);
}

}  // extern "C"
