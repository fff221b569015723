// Host build of the device field/curve headers (their plain-C twins of the PTX carry chains), so the
// limb logic is unit-tested on the CPU box.  Built by tests/test_ff_host.py; not part of the product.
#include "../tiny-ram-halo2_b200/csrc/ff.cuh"
#include "../tiny-ram-halo2_b200/csrc/ec.cuh"
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
extern "C" {
void ffh_op(int field, int o, const uint32_t* a, const uint32_t* b, uint32_t* r, size_t n) {
  if (field == 0) op<FpParams>(o, a, b, r, n); else op<FqParams>(o, a, b, r, n);
}
void ffh_ecop(int base_field, const uint32_t* pts, const int* neg, size_t n, const uint32_t* q_xyzz, int dbls, uint32_t* out_aff) {
  if (base_field == 0) ecop<FpParams>(pts, neg, n, q_xyzz, dbls, out_aff); else ecop<FqParams>(pts, neg, n, q_xyzz, dbls, out_aff);
}
}
