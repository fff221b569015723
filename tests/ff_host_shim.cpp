// Host build of the device field/curve headers (their plain-C twins of the PTX carry chains), so the
// limb logic is unit-tested on the CPU box.  Built by tests/test_ff_host.py; not part of the product.
#include "../tiny-ram-halo2_b200/csrc/ff.cuh"
#include "../tiny-ram-halo2_b200/csrc/ec.cuh"
#include "../tiny-ram-halo2_b200/csrc/bucket_reduce.cuh"
#include <vector>
#include <cstring>
using namespace ff;

template <class PR> static void op(int o, const uint32_t* a, const uint32_t* b, uint32_t* r, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    Fe<PR> x, y, z;
    memcpy(x.v, a + 8 * i, 32); memcpy(y.v, b + 8 * i, 32);
    switch (o) {
      case 0: z = fe_add(x, y); break;
      case 1: z = fe_sub(x, y); break;
      case 2: z = fe_mul(x, y); break;
      case 3: z = fe_inv(x); break;
      case 4: z = fe_sqr(x); break;
      case 5: z = fe_from_mont(x); break;
      case 6: z = fe_to_mont(x); break;
      default: z = fe_neg(x);
    }
    memcpy(r + 8 * i, z.v, 32);
  }
}
// xyzz accumulate: out = sum_i sign_i * P_i done with mixed adds, then + Q (full add), then doubled k times
template <class PR> static void ecop(const uint32_t* pts, const int* neg, size_t n, const uint32_t* q_xyzz, int dbls, uint32_t* out_aff) {
  ec::XYZZ<PR> acc = ec::xyzz_identity<PR>();
  for (size_t i = 0; i < n; ++i) {
    ec::Affine<PR> p; memcpy(p.x.v, pts + 16 * i, 32); memcpy(p.y.v, pts + 16 * i + 8, 32);
    if (neg[i]) p.y = fe_neg(p.y);
    ec::xyzz_add_mixed(acc, p);
  }
  if (q_xyzz) {
    ec::XYZZ<PR> q; memcpy(&q, q_xyzz, 128);
    ec::xyzz_add(acc, q);
  }
  for (int i = 0; i < dbls; ++i) ec::xyzz_dbl(acc);
  ec::Affine<PR> a = ec::xyzz_to_affine(acc);
  memcpy(out_aff, a.x.v, 32); memcpy(out_aff + 8, a.y.v, 32);
}

// the two-level weighted bucket sum of bucket_reduce.cuh the way msm_reduce_chunks2 / msm_reduce_sets2 run it: chunks of
// 2^log_chunk buckets, `threads` leaves of 2^log_m chunks each, a binary tree over the leaves
template <class PR> static void bucket_reduce2(const uint32_t* buckets_aff, unsigned threads, unsigned log_m, unsigned log_chunk, uint32_t* out_aff) {
  const unsigned S = 1u << log_chunk, m = 1u << log_m, chunks = threads * m;
  auto bucket = [&](size_t b) {
    ec::Affine<PR> p; memcpy(p.x.v, buckets_aff + 16 * b, 32); memcpy(p.y.v, buckets_aff + 16 * b + 8, 32);
    return ec::xyzz_from_affine(p);
  };
  std::vector<ec::XYZZ<PR>> acc(chunks), tot(chunks);
  for (unsigned ch = 0; ch < chunks; ++ch) {
    ec::XYZZ<PR> run = ec::xyzz_identity<PR>(), a = ec::xyzz_identity<PR>();
    for (int k = (int)S - 1; k >= 0; --k) { ec::xyzz_add(run, bucket((size_t)ch * S + k)); ec::xyzz_add(a, run); }
    acc[ch] = a; tot[ch] = run;
  }
  std::vector<ec::WNode<PR>> node(threads);
  for (unsigned t = 0; t < threads; ++t)
    node[t] = ec::wnode_leaf<PR>(m, [&](unsigned ch) { return acc[t * m + ch]; }, [&](unsigned ch) { return tot[t * m + ch]; });
  unsigned lv = 0;
  for (unsigned stride = 1; stride < threads; stride <<= 1, ++lv)
    for (unsigned t = 0; t < threads; t += 2 * stride) ec::wnode_combine(node[t], node[t + stride], log_m + lv);
  ec::Affine<PR> r = ec::xyzz_to_affine(ec::wnode_root(node[0], log_chunk));
  memcpy(out_aff, r.x.v, 32); memcpy(out_aff + 8, r.y.v, 32);
}

extern "C" {
void ffh_bucket_reduce2(int base_field, const uint32_t* buckets_aff, unsigned threads, unsigned log_m, unsigned log_chunk, uint32_t* out_aff) {
  if (base_field == 0) bucket_reduce2<FpParams>(buckets_aff, threads, log_m, log_chunk, out_aff);
  else bucket_reduce2<FqParams>(buckets_aff, threads, log_m, log_chunk, out_aff);
}
void ffh_op(int field, int o, const uint32_t* a, const uint32_t* b, uint32_t* r, size_t n) {
  if (field == 0) op<FpParams>(o, a, b, r, n); else op<FqParams>(o, a, b, r, n);
}
void ffh_ecop(int base_field, const uint32_t* pts, const int* neg, size_t n, const uint32_t* q_xyzz, int dbls, uint32_t* out_aff) {
  if (base_field == 0) ecop<FpParams>(pts, neg, n, q_xyzz, dbls, out_aff); else ecop<FqParams>(pts, neg, n, q_xyzz, dbls, out_aff);
}
}
