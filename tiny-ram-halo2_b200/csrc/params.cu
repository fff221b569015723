// halo2_proofs::poly::commitment::Params::new(k) on the GPU.
//
// Reached from the reference at /root/reference/src/test_utils.rs:21,89 (`Params::<EqAffine>::new(k)`); the code it runs is
// in the un-vendored crates halo2_proofs 0.2.0 (poly/commitment.rs) and pasta_curves 0.4.1 (hashtocurve.rs, curves.rs):
//   * CurveExt::hash_to_curve("Halo2-Parameters")([0] ++ u32_le(i))      -> h2c_kernel  (one thread per generator:
//       expand_message_xmd over BLAKE2b-512, simplified SWU onto the iso-curve, affine add, 3-isogeny)
//   * best_fft over Vec<C::Curve> with alpha^-1 (the group iFFT g -> g_lagrange), then * TWO_INV^k, batch_normalize
//       -> gfft_* kernels (bit-reversed load to XYZZ, one kernel per radix-2 stage with a fixed 4-bit-window scalar
//          multiplication by the twiddle, final scaling, one batched inversion to return to affine)
// Group elements are unique, so the coordinate system and the window method are invisible in the affine results.
#include <string.h>

#include "common.cuh"
#include "ec.cuh"
#include "h2c.cuh"

using namespace ff;
using namespace ec;
using h2c::H2cConsts;

namespace {

template <class PR>
__global__ void __launch_bounds__(64) h2c_kernel(h2c::H2cConsts K, const uint8_t* d_msgs, size_t n, uint4* out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fe<PR> x, y;
  h2c::h2c_point<PR>(K, d_msgs, i, x, y);
  fe_store(out + 4 * i, x);
  fe_store(out + 4 * i + 2, y);
}

// ---- bulk random field elements: a CSPRNG seed expanded on the device ---------------------------------------------------------------
// out[i] = Field::random over the 64 bytes BLAKE2b-512(key ++ u64_le(first + i)) -- pasta_curves' from_u512 (lo * R^2 + hi * R^3 in
// Montgomery products) -- so a random polynomial (vanishing::Argument::commit's random poly, the IPA's S) costs the caller's RNG
// 32 bytes instead of 64 n, on every GPU alike.
struct Key32 { uint8_t b[32]; };
template <class PR>
__global__ void __launch_bounds__(128) random_field_kernel(Key32 key, uint64_t first, size_t n, uint4* out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  h2c::B2b s;
  h2c::b2b_init(s);
  h2c::b2b_update(s, key.b, 32);
  uint8_t ctr[8], d[64];
  const uint64_t c = first + i;
  for (int k = 0; k < 8; ++k) ctr[k] = (uint8_t)(c >> (8 * k));
  h2c::b2b_update(s, ctr, 8);
  h2c::b2b_final(s, d);
  Fe<PR> lo, hi, r2;
  for (int k = 0; k < 8; ++k) {
    lo.v[k] = (uint32_t)d[4 * k] | ((uint32_t)d[4 * k + 1] << 8) | ((uint32_t)d[4 * k + 2] << 16) | ((uint32_t)d[4 * k + 3] << 24);
    hi.v[k] = (uint32_t)d[32 + 4 * k] | ((uint32_t)d[32 + 4 * k + 1] << 8) | ((uint32_t)d[32 + 4 * k + 2] << 16) | ((uint32_t)d[32 + 4 * k + 3] << 24);
    r2.v[k] = PR::r2(k);
  }
  const Fe<PR> r3 = fe_mul(r2, r2);
  for (int k = 0; k < 3; ++k) { fe_final_sub(lo); fe_final_sub(hi); }      // halves < 2^256 < 4p: fe_mul expects reduced operands
  fe_store(out + 2 * i, fe_add(fe_mul(lo, r2), fe_mul(hi, r3)));
}

// ---- group FFT ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ size_t bitrev(size_t x, unsigned bits) { return bits ? (size_t)(__brevll((unsigned long long)x) >> (64 - bits)) : 0; }

template <class BPR> __device__ __forceinline__ XYZZ<BPR> xyzz_load(const uint4* p) {
  XYZZ<BPR> r; r.x = fe_load<BPR>(p); r.y = fe_load<BPR>(p + 2); r.zz = fe_load<BPR>(p + 4); r.zzz = fe_load<BPR>(p + 6); return r;
}
template <class BPR> __device__ __forceinline__ void xyzz_store(uint4* p, const XYZZ<BPR>& r) {
  fe_store(p, r.x); fe_store(p + 2, r.y); fe_store(p + 4, r.zz); fe_store(p + 6, r.zzz);
}

// acc[i] = affine[bitrev(i)] as XYZZ
template <class BPR>
__global__ void gfft_load_kernel(const uint4* aff, unsigned log_n, uint4* acc) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >> log_n) return;
  const size_t j = bitrev(i, log_n);
  Affine<BPR> p; p.x = fe_load<BPR>(aff + 4 * j); p.y = fe_load<BPR>(aff + 4 * j + 2);
  xyzz_store<BPR>(acc + 8 * i, xyzz_from_affine(p));
}

// [k] P for the canonical 256-bit integer k (8 x u32), fixed 4-bit windows: the same instruction stream in every lane
template <class BPR>
__device__ XYZZ<BPR> xyzz_mul_w4(const XYZZ<BPR>& p, const uint32_t k[8]) {
  XYZZ<BPR> tab[16];
  tab[0] = xyzz_identity<BPR>();
  tab[1] = p;
  for (int i = 2; i < 16; ++i) { tab[i] = tab[i - 1]; xyzz_add(tab[i], p); }   // xyzz_add doubles when the operands coincide
  XYZZ<BPR> r = xyzz_identity<BPR>();
  for (int w = 63; w >= 0; --w) {
    xyzz_dbl(r); xyzz_dbl(r); xyzz_dbl(r); xyzz_dbl(r);
    const uint32_t d = (k[w >> 3] >> ((w & 7) * 4)) & 15u;
    if (d) xyzz_add(r, tab[d]);
  }
  return r;
}

// one radix-2 DIT stage over bit-reversed data: pairs (s + i, s + i + half), twiddle tw[i * (n / 2 / half)]
template <class SPR, class BPR>
__global__ void __launch_bounds__(128) gfft_stage_kernel(uint4* acc, unsigned log_n, unsigned stage, const uint4* tw) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >> (log_n - 1)) return;
  const size_t half = (size_t)1 << stage;
  const size_t i = t & (half - 1);
  const size_t lo = ((t >> stage) << (stage + 1)) + i, hi = lo + half;
  XYZZ<BPR> b = xyzz_load<BPR>(acc + 8 * hi);
  if (i != 0) {
    Fe<SPR> w = fe_from_mont(fe_load_ro<SPR>(tw + 2 * (i << (log_n - 1 - stage))));
    b = xyzz_mul_w4<BPR>(b, w.v);
  }
  XYZZ<BPR> a = xyzz_load<BPR>(acc + 8 * lo);
  XYZZ<BPR> s = a;
  xyzz_add(s, b);
  xyzz_add(a, xyzz_neg(b));
  xyzz_store<BPR>(acc + 8 * lo, s);
  xyzz_store<BPR>(acc + 8 * hi, a);
}

// acc[i] = [scale] acc[i] (when has_scale), den[i] = ZZ * ZZZ
struct ScaleBits { uint32_t v[8]; int on; };
template <class BPR>
__global__ void __launch_bounds__(128) gfft_scale_kernel(uint4* acc, size_t n, ScaleBits sc, uint4* den) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  XYZZ<BPR> p = xyzz_load<BPR>(acc + 8 * i);
  if (sc.on) { p = xyzz_mul_w4<BPR>(p, sc.v); xyzz_store<BPR>(acc + 8 * i, p); }
  fe_store(den + 2 * i, fe_mul(p.zz, p.zzz));
}
template <class BPR>
__global__ void gfft_norm_kernel(const uint4* acc, const uint4* inv, size_t n, uint4* out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4* a = acc + 8 * i;
  Fe<BPR> zz = fe_load<BPR>(a + 4);
  Fe<BPR> x = fe_zero<BPR>(), y = fe_zero<BPR>();
  if (!fe_is_zero(zz)) {
    Fe<BPR> iv = fe_load<BPR>(inv + 2 * i);
    x = fe_mul(fe_load<BPR>(a), fe_mul(iv, fe_load<BPR>(a + 6)));
    y = fe_mul(fe_load<BPR>(a + 2), fe_mul(iv, zz));
  }
  fe_store(out + 4 * i, x); fe_store(out + 4 * i + 2, y);
}

// ---- host side ------------------------------------------------------------------------------------------------------------------
template <class PR> Fe<PR> fe_from_limbs64(const uint64_t* l) {
  Fe<PR> r;
  for (int i = 0; i < 4; ++i) { r.v[2 * i] = (uint32_t)l[i]; r.v[2 * i + 1] = (uint32_t)(l[i] >> 32); }
  return r;
}

int make_consts(trp_ctx* ctx, const char* domain_prefix, H2cConsts& K) {
  if (!h2c::make_consts(ctx->curve == TRP_CURVE_PALLAS, domain_prefix, K))
    TRP_FAIL(ctx, TRP_E_INVALID, "domain prefix too long (pasta_curves asserts 22 + len(curve_id) + len(prefix) < 256)");
  return TRP_OK;
}

int h2c_launch(trp_ctx* ctx, const H2cConsts& K, const uint8_t* d_msgs, size_t n, void* d_out) {
  if (n == 0) return TRP_OK;
  const unsigned blocks = (unsigned)((n + 63) / 64);
  if (base_field_of(ctx->curve) == 0) h2c_kernel<FpParams><<<blocks, 64, 0, ctx->stream>>>(K, d_msgs, n, (uint4*)d_out);
  else h2c_kernel<FqParams><<<blocks, 64, 0, ctx->stream>>>(K, d_msgs, n, (uint4*)d_out);
  TRP_LAUNCHED(ctx);
  return TRP_OK;
}

template <class SPR, class BPR>
int group_fft_run(trp_ctx* ctx, void* d_points, unsigned log_n, const uint64_t omega[4], const uint64_t* scale, void* ws) {
  const size_t n = (size_t)1 << log_n;
  uint4* acc = (uint4*)ws;
  uint4* den = (uint4*)((char*)ws + ws_align(n * 128));
  const unsigned blocks = (unsigned)((n + 127) / 128);
  gfft_load_kernel<BPR><<<blocks, 128, 0, ctx->stream>>>((const uint4*)d_points, log_n, acc);
  TRP_LAUNCHED(ctx);
  if (log_n > 0) {
    const void* tw = nullptr;
    TRP_TRY(trp_get_powers(ctx, scalar_field_of(ctx->curve), log_n, omega, &tw));
    const unsigned hb = (unsigned)((n / 2 + 127) / 128);
    for (unsigned s = 0; s < log_n; ++s) {
      gfft_stage_kernel<SPR, BPR><<<hb, 128, 0, ctx->stream>>>(acc, log_n, s, (const uint4*)tw);
      TRP_LAUNCHED(ctx);
    }
  }
  ScaleBits sb; sb.on = 0;
  if (scale) {
    Fe<SPR> c = fe_from_mont(fe_from_limbs64<SPR>(scale));
    for (int i = 0; i < 8; ++i) sb.v[i] = c.v[i];
    sb.on = 1;
  }
  gfft_scale_kernel<BPR><<<blocks, 128, 0, ctx->stream>>>(acc, n, sb, den);
  TRP_LAUNCHED(ctx);
  TRP_TRY(trp_batch_invert_impl(ctx, base_field_of(ctx->curve), den, nullptr, den, n));
  gfft_norm_kernel<BPR><<<blocks, 128, 0, ctx->stream>>>(acc, den, n, (uint4*)d_points);
  TRP_LAUNCHED(ctx);
  return TRP_OK;
}

}  // namespace

size_t trp_group_fft_ws_bytes(unsigned log_n) {
  const size_t n = (size_t)1 << log_n;
  return ws_align(n * 128) + ws_align(n * 32) + 256;
}

int trp_group_fft_impl(trp_ctx* ctx, void* d_points, unsigned log_n, const uint64_t omega[4], const uint64_t* scale, void* ws) {
  if (ctx->curve == TRP_CURVE_PALLAS) return group_fft_run<FqParams, FpParams>(ctx, d_points, log_n, omega, scale, ws);
  return group_fft_run<FpParams, FqParams>(ctx, d_points, log_n, omega, scale, ws);
}

int trp_random_field_impl(trp_ctx* ctx, int field, const uint8_t key[32], uint64_t first, size_t n, void* d_out) {
  if (n == 0) return TRP_OK;
  Key32 k;
  memcpy(k.b, key, 32);
  unsigned blocks = (unsigned)((n + 127) / 128);
  if (field == 0) random_field_kernel<FpParams><<<blocks, 128, 0, ctx->stream>>>(k, first, n, (uint4*)d_out);
  else random_field_kernel<FqParams><<<blocks, 128, 0, ctx->stream>>>(k, first, n, (uint4*)d_out);
  TRP_LAUNCHED(ctx);
  return TRP_OK;
}

int trp_hash_to_curve_impl(trp_ctx* ctx, const char* domain_prefix, const uint8_t* d_msgs, size_t msg_len, const uint8_t* msg_prefix,
                           size_t prefix_len, int append_index, uint64_t first_index, size_t n, void* d_out) {
  H2cConsts K;
  TRP_TRY(make_consts(ctx, domain_prefix, K));
  if (prefix_len > sizeof(K.msg_prefix)) TRP_FAIL(ctx, TRP_E_INVALID, "message prefix longer than %zu bytes", sizeof(K.msg_prefix));
  if (prefix_len) memcpy(K.msg_prefix, msg_prefix, prefix_len);
  K.prefix_len = (uint32_t)prefix_len;
  K.append_index = append_index ? 1 : 0;
  K.first_index = first_index;
  K.msg_len = (uint32_t)msg_len;
  return h2c_launch(ctx, K, d_msgs, n, d_out);
}
