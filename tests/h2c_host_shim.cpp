// Host build of csrc/h2c.cuh (pasta_curves' hash-to-curve as host/device functions) so that the byte and limb logic is
// unit-tested on the CPU box against oracle/params_model.py.  Built by tests/test_ff_host.py; not part of the product.
#include "../tiny-ram-halo2_b200/csrc/h2c.cuh"
using namespace ff;

extern "C" {
// out: n x 16 u32 (x, y Montgomery); returns 0 on success
int h2ch_hash(int pallas, const char* domain_prefix, const uint8_t* msgs, size_t msg_len, size_t n, uint32_t* out) {
  h2c::H2cConsts K;
  if (!h2c::make_consts(pallas != 0, domain_prefix, K)) return -1;
  K.msg_len = (uint32_t)msg_len;
  for (size_t i = 0; i < n; ++i) {
    if (pallas) { Fe<FpParams> x, y; h2c::h2c_point<FpParams>(K, msgs, i, x, y); memcpy(out + 16 * i, x.v, 32); memcpy(out + 16 * i + 8, y.v, 32); }
    else { Fe<FqParams> x, y; h2c::h2c_point<FqParams>(K, msgs, i, x, y); memcpy(out + 16 * i, x.v, 32); memcpy(out + 16 * i + 8, y.v, 32); }
  }
  return 0;
}
// pieces, for narrowing down a mismatch: BLAKE2b-512 of a buffer; from_bytes_wide of a big-endian digest; one SWU map; sqrt
void h2ch_blake2b(const uint8_t* data, size_t len, uint8_t out[64]) {
  h2c::B2b s; h2c::b2b_init(s); h2c::b2b_update(s, data, (uint32_t)len); h2c::b2b_final(s, out);
}
void h2ch_from_be64(int pallas, const uint8_t d[64], uint32_t out[8]) {
  h2c::H2cConsts K; h2c::make_consts(pallas != 0, "", K);
  if (pallas) { Fe<FpParams> r = h2c::fe_from_be64<FpParams>(K, d); memcpy(out, r.v, 32); }
  else { Fe<FqParams> r = h2c::fe_from_be64<FqParams>(K, d); memcpy(out, r.v, 32); }
}
void h2ch_swu(int pallas, const uint32_t u[8], uint32_t out[16]) {
  h2c::H2cConsts K; h2c::make_consts(pallas != 0, "", K);
  if (pallas) { Fe<FpParams> a, x, y; memcpy(a.v, u, 32); h2c::swu_map<FpParams>(K, a, x, y); memcpy(out, x.v, 32); memcpy(out + 8, y.v, 32); }
  else { Fe<FqParams> a, x, y; memcpy(a.v, u, 32); h2c::swu_map<FqParams>(K, a, x, y); memcpy(out, x.v, 32); memcpy(out + 8, y.v, 32); }
}
int h2ch_sqrt(int pallas, const uint32_t a[8], uint32_t out[8]) {
  h2c::H2cConsts K; h2c::make_consts(pallas != 0, "", K);
  bool ok;
  if (pallas) { Fe<FpParams> x, r; memcpy(x.v, a, 32); ok = h2c::fe_sqrt<FpParams>(K, x, r); memcpy(out, r.v, 32); }
  else { Fe<FqParams> x, r; memcpy(x.v, a, 32); ok = h2c::fe_sqrt<FqParams>(K, x, r); memcpy(out, r.v, 32); }
  return ok ? 1 : 0;
}
}
