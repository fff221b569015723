// pasta_curves 0.4.1 hashtocurve.rs (hash_to_field, map_to_curve_simple_swu, iso_map) and CurveExt::hash_to_curve as
// host/device functions: the device build is params.cu's h2c_kernel, the host build (tests/h2c_host_shim.cpp) lets the byte
// and limb logic be unit-tested on the CPU box against oracle/params_model.py.
#pragma once
#include <string.h>

#include "ff.cuh"

#if defined(__CUDACC__)
#define H2C_FN __host__ __device__ inline
#else
#define H2C_FN inline
#endif

namespace h2c {
using namespace ff;

// ---- BLAKE2b-512, unkeyed, empty personalisation (blake2b_simd::Params::new().hash_length(64).personal(&[0; 16])) ---------
H2C_FN uint64_t b2b_iv(int i) {
  const uint64_t iv[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                          0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
  return iv[i];
}

struct B2b {
  uint64_t h[8];
  uint64_t t;
  uint8_t buf[128];
  uint32_t len;
};

H2C_FN uint64_t rotr64(uint64_t x, int r) { return (x >> r) | (x << (64 - r)); }

H2C_FN void b2b_compress(B2b& s, bool last) {
  const uint8_t sigma[12][16] = {
      {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
      {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
      {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
      {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
      {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
      {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
  uint64_t m[16], v[16];
  for (int i = 0; i < 16; ++i) {
    uint64_t w = 0;
    for (int b = 7; b >= 0; --b) w = (w << 8) | s.buf[8 * i + b];
    m[i] = w;
  }
  for (int i = 0; i < 8; ++i) { v[i] = s.h[i]; v[i + 8] = b2b_iv(i); }
  v[12] ^= s.t;
  if (last) v[14] = ~v[14];
#define B2B_G(a, b, c, d, x, y)              \
  v[a] = v[a] + v[b] + (x); v[d] = rotr64(v[d] ^ v[a], 32); \
  v[c] = v[c] + v[d];       v[b] = rotr64(v[b] ^ v[c], 24); \
  v[a] = v[a] + v[b] + (y); v[d] = rotr64(v[d] ^ v[a], 16); \
  v[c] = v[c] + v[d];       v[b] = rotr64(v[b] ^ v[c], 63);
  for (int r = 0; r < 12; ++r) {
    const uint8_t* g = sigma[r];
    B2B_G(0, 4, 8, 12, m[g[0]], m[g[1]])
    B2B_G(1, 5, 9, 13, m[g[2]], m[g[3]])
    B2B_G(2, 6, 10, 14, m[g[4]], m[g[5]])
    B2B_G(3, 7, 11, 15, m[g[6]], m[g[7]])
    B2B_G(0, 5, 10, 15, m[g[8]], m[g[9]])
    B2B_G(1, 6, 11, 12, m[g[10]], m[g[11]])
    B2B_G(2, 7, 8, 13, m[g[12]], m[g[13]])
    B2B_G(3, 4, 9, 14, m[g[14]], m[g[15]])
  }
#undef B2B_G
  for (int i = 0; i < 8; ++i) s.h[i] ^= v[i] ^ v[i + 8];
}

H2C_FN void b2b_init(B2b& s) {
  for (int i = 0; i < 8; ++i) s.h[i] = b2b_iv(i);
  s.h[0] ^= 0x01010040ull;   // digest 64 bytes, no key, fanout = depth = 1
  s.t = 0; s.len = 0;
}
H2C_FN void b2b_update(B2b& s, const uint8_t* p, uint32_t n) {
  for (uint32_t i = 0; i < n; ++i) {
    if (s.len == 128) { s.t += 128; b2b_compress(s, false); s.len = 0; }
    s.buf[s.len++] = p[i];
  }
}
H2C_FN void b2b_final(B2b& s, uint8_t out[64]) {
  s.t += s.len;
  for (uint32_t i = s.len; i < 128; ++i) s.buf[i] = 0;
  b2b_compress(s, true);
  for (int i = 0; i < 64; ++i) out[i] = (uint8_t)(s.h[i >> 3] >> (8 * (i & 7)));
}

// ---- constants of one curve's hash-to-curve suite, prepared on the host (Montgomery form over the BASE field) ----------------
struct H2cConsts {
  uint32_t a[8], b[8], z[8];       // iso-curve y^2 = x^3 + a x + b, SWU parameter Z = -13
  uint32_t iso[13][8];             // pasta_curves' ISOGENY_CONSTANTS
  uint32_t r3[8];                  // 2^768 mod p: from_bytes_wide = lo * R2 + hi * R3 (Montgomery products)
  uint32_t root[8];                // ROOT_OF_UNITY (order 2^32)
  uint32_t t_minus1_over2[8];      // (T - 1) / 2, p - 1 = 2^32 * T, plain integer limbs
  uint8_t dst_prime[256];          // domain_prefix "-" curve_id "_XMD:BLAKE2b_SSWU_RO_" ++ [len]
  uint32_t dst_len;
  uint8_t msg_prefix[64];
  uint32_t prefix_len;
  uint32_t append_index;           // message = msg_prefix ++ u32_le(first_index + i)
  uint64_t first_index;
  uint32_t msg_len;                // when d_msgs != NULL: fixed message length
};

template <class PR> H2C_FN Fe<PR> fe_of(const uint32_t* l) {
  Fe<PR> r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = l[i];
  return r;
}

// Tonelli-Shanks with p - 1 = 2^32 * T; false for a non-residue
template <class PR>
H2C_FN bool fe_sqrt(const H2cConsts& K, const Fe<PR>& a, Fe<PR>& out) {
  if (fe_is_zero(a)) { out = a; return true; }
  Fe<PR> w = fe_pow(a, K.t_minus1_over2, 8);
  Fe<PR> r = fe_mul(a, w);          // a^((T+1)/2)
  Fe<PR> t = fe_mul(r, w);          // a^T
  Fe<PR> c = fe_of<PR>(K.root);
  const Fe<PR> one = fe_one<PR>();
  int m = 32;
  while (!fe_eq(t, one)) {
    int i = 0;
    Fe<PR> t2 = t;
    while (!fe_eq(t2, one) && i < m) { t2 = fe_sqr(t2); ++i; }
    if (i >= m) return false;
    Fe<PR> bb = c;
    for (int j = 0; j < m - i - 1; ++j) bb = fe_sqr(bb);
    m = i;
    c = fe_sqr(bb);
    t = fe_mul(t, c);
    r = fe_mul(r, bb);
  }
  out = r;
  return true;
}

template <class PR> H2C_FN bool fe_is_odd(const Fe<PR>& a) { return fe_from_mont(a).v[0] & 1; }

// 64-byte digest read BIG-endian, reduced: pasta's from_bytes_wide on the reversed bytes
template <class PR>
H2C_FN Fe<PR> fe_from_be64(const H2cConsts& K, const uint8_t d[64]) {
  Fe<PR> lo, hi;
  for (int j = 0; j < 16; ++j) {
    uint32_t w = (uint32_t)d[63 - 4 * j] | ((uint32_t)d[62 - 4 * j] << 8) | ((uint32_t)d[61 - 4 * j] << 16) | ((uint32_t)d[60 - 4 * j] << 24);
    if (j < 8) lo.v[j] = w; else hi.v[j - 8] = w;
  }
  Fe<PR> r2;
#pragma unroll
  for (int i = 0; i < 8; ++i) r2.v[i] = PR::r2(i);
  // the halves are below 2^256 < 4p: bring them below p first (fe_mul expects reduced operands)
  for (int i = 0; i < 3; ++i) { fe_final_sub(lo); fe_final_sub(hi); }
  return fe_add(fe_mul(lo, r2), fe_mul(hi, fe_of<PR>(K.r3)));
}

template <class PR>
H2C_FN void swu_map(const H2cConsts& K, const Fe<PR>& u, Fe<PR>& x, Fe<PR>& y) {
  const Fe<PR> a = fe_of<PR>(K.a), b = fe_of<PR>(K.b), z = fe_of<PR>(K.z);
  Fe<PR> z_u2 = fe_mul(z, fe_sqr(u));
  Fe<PR> ta = fe_add(fe_sqr(z_u2), z_u2);
  Fe<PR> num_x1 = fe_mul(b, fe_add(ta, fe_one<PR>()));
  Fe<PR> div = fe_mul(a, fe_is_zero(ta) ? z : fe_neg(ta));
  Fe<PR> x1 = fe_mul(num_x1, fe_inv(div));
  Fe<PR> gx1 = fe_add(fe_mul(fe_add(fe_sqr(x1), a), x1), b);
  Fe<PR> yy;
  if (fe_sqrt(K, gx1, yy)) {
    x = x1;
  } else {
    x = fe_mul(z_u2, x1);
    Fe<PR> gx2 = fe_add(fe_mul(fe_add(fe_sqr(x), a), x), b);
    fe_sqrt(K, gx2, yy);      // exactly one of gx1, gx2 is a square
  }
  if (fe_is_odd(u) != fe_is_odd(yy)) yy = fe_neg(yy);
  y = yy;
}

// message i of a call: d_msgs + i * msg_len when d_msgs != NULL, else msg_prefix ++ u32_le(first_index + i)
template <class PR>
H2C_FN void h2c_point(const H2cConsts& K, const uint8_t* d_msgs, size_t i, Fe<PR>& ox, Fe<PR>& oy) {
  // ---- hash_to_field: expand_message_xmd, 2 x 64 bytes --------------------------------------------------------------------
  uint8_t b0[64], b1[64], b2[64];
  B2b s;
  b2b_init(s);
  { uint8_t zero[128]; for (int j = 0; j < 128; ++j) zero[j] = 0; b2b_update(s, zero, 128); }
  if (d_msgs) {
    b2b_update(s, d_msgs + i * K.msg_len, K.msg_len);
  } else {
    b2b_update(s, K.msg_prefix, K.prefix_len);
    if (K.append_index) {
      uint32_t idx = (uint32_t)(K.first_index + i);
      uint8_t le[4] = {(uint8_t)idx, (uint8_t)(idx >> 8), (uint8_t)(idx >> 16), (uint8_t)(idx >> 24)};
      b2b_update(s, le, 4);
    }
  }
  { uint8_t l[3] = {0, 128, 0}; b2b_update(s, l, 3); }
  b2b_update(s, K.dst_prime, K.dst_len);
  b2b_final(s, b0);
  b2b_init(s);
  b2b_update(s, b0, 64);
  { uint8_t c = 1; b2b_update(s, &c, 1); }
  b2b_update(s, K.dst_prime, K.dst_len);
  b2b_final(s, b1);
  for (int j = 0; j < 64; ++j) b2[j] = b0[j] ^ b1[j];
  b2b_init(s);
  b2b_update(s, b2, 64);
  { uint8_t c = 2; b2b_update(s, &c, 1); }
  b2b_update(s, K.dst_prime, K.dst_len);
  b2b_final(s, b2);
  Fe<PR> u0 = fe_from_be64<PR>(K, b1), u1 = fe_from_be64<PR>(K, b2);
  // ---- two SWU maps onto the iso-curve, affine sum ------------------------------------------------------------------------------
  Fe<PR> x0, y0, x1, y1;
  swu_map<PR>(K, u0, x0, y0);
  swu_map<PR>(K, u1, x1, y1);
  const Fe<PR> a = fe_of<PR>(K.a);
  ox = fe_zero<PR>(); oy = fe_zero<PR>();
  bool inf = false;
  Fe<PR> lam_num, lam_den;
  if (fe_eq(x0, x1)) {
    if (fe_is_zero(fe_add(y0, y1))) inf = true;
    else { Fe<PR> xx = fe_sqr(x0); lam_num = fe_add(fe_add(fe_dbl(xx), xx), a); lam_den = fe_dbl(y0); }
  } else {
    lam_num = fe_sub(y1, y0); lam_den = fe_sub(x1, x0);
  }
  if (!inf) {
    Fe<PR> lam = fe_mul(lam_num, fe_inv(lam_den));
    Fe<PR> x3 = fe_sub(fe_sub(fe_sqr(lam), x0), x1);
    Fe<PR> y3 = fe_sub(fe_mul(lam, fe_sub(x0, x3)), y0);
    // ---- iso_map: the 3-isogeny onto y^2 = x^3 + 5 ------------------------------------------------------------------------------
#define ISO(j) fe_of<PR>(K.iso[j])
    Fe<PR> num_x = fe_add(fe_mul(fe_add(fe_mul(fe_add(fe_mul(ISO(0), x3), ISO(1)), x3), ISO(2)), x3), ISO(3));
    Fe<PR> div_x = fe_add(fe_mul(fe_add(x3, ISO(4)), x3), ISO(5));
    Fe<PR> num_y = fe_mul(fe_add(fe_mul(fe_add(fe_mul(fe_add(fe_mul(ISO(6), x3), ISO(7)), x3), ISO(8)), x3), ISO(9)), y3);
    Fe<PR> div_y = fe_add(fe_mul(fe_add(fe_mul(fe_add(x3, ISO(10)), x3), ISO(11)), x3), ISO(12));
#undef ISO
    Fe<PR> dd = fe_mul(div_x, div_y);
    if (!fe_is_zero(dd)) {                    // dd = 0: the point lies in the isogeny's kernel and maps to the identity
      Fe<PR> inv = fe_inv(dd);
      ox = fe_mul(num_x, fe_mul(inv, div_y));
      oy = fe_mul(num_y, fe_mul(inv, div_x));
    }
  }
}


// ---- host side: the constants of a suite ---------------------------------------------------------------------------------------
template <class PR> inline Fe<PR> fe_from_limbs64(const uint64_t* l) {
  Fe<PR> r;
  for (int i = 0; i < 4; ++i) { r.v[2 * i] = (uint32_t)l[i]; r.v[2 * i + 1] = (uint32_t)(l[i] >> 32); }
  return r;
}
struct CurveSuite { const char* id; uint64_t a[4]; uint64_t root[4]; uint64_t iso[13][4]; };
// canonical values.  a: pasta_curves IsoEp / IsoEq; iso: the 3-isogeny derived by Velu's formulas in oracle/params_model.py (the
// Pallas set equals pasta_curves' published ISOGENY_CONSTANTS); root: ROOT_OF_UNITY of the base field.
static const CurveSuite SUITE_PALLAS = {
    "pallas",
    {0x92bb4b0b657a014bull, 0xb74134581a27a59full, 0x49be2d7258370742ull, 0x18354a2eb0ea8c9cull},
    {0xbdad6fabd87ea32full, 0xea322bf2b7bb7584ull, 0x362120830561f81aull, 0x2bce74deac30ebdaull},
    {{0x775f6034aaaaaaabull, 0x4081775473d8375bull, 0xe38e38e38e38e38eull, 0x0e38e38e38e38e38ull},
     {0x8cf863b02814fb76ull, 0x0f93b82ee4b99495ull, 0x267c7ffa51cf412aull, 0x3509afd51872d88eull},
     {0x0eb64faef37ea4f7ull, 0x380af066cfeb6d69ull, 0x98c7d7ac3d98fd13ull, 0x17329b9ec5253753ull},
     {0xeebec06955555580ull, 0x8102eea8e7b06eb6ull, 0xc71c71c71c71c71cull, 0x1c71c71c71c71c71ull},
     {0xc47f2ab668bcd71full, 0x9c434ac1c96b6980ull, 0x5a607fcce0494a79ull, 0x1d572e7ddc099cffull},
     {0x2aa3af1eae5b6604ull, 0xb4abf9fb9a1fc81cull, 0x1d13bf2a7f22b105ull, 0x325669becaecd5d1ull},
     {0x5ad985b5e38e38e4ull, 0x7642b01ad461bad2ull, 0x4bda12f684bda12full, 0x1a12f684bda12f68ull},
     {0xc67c31d8140a7dbbull, 0x07c9dc17725cca4aull, 0x133e3ffd28e7a095ull, 0x1a84d7ea8c396c47ull},
     {0x02e2be87d225b234ull, 0x1765e924f7459378ull, 0x303216cce1db9ff1ull, 0x3fb98ff0d2ddcaddull},
     {0x93e53ab371c71c4full, 0x0ac03e8e134eb3e4ull, 0x7b425ed097b425edull, 0x025ed097b425ed09ull},
     {0x5a28279b1d1b42aeull, 0x5941a3a4a97aa1b3ull, 0x0790bfb3506defb6ull, 0x0c02c5bcca0e6b7full},
     {0x4d90ab820b12320aull, 0xd976bbfabbc5661dull, 0x573b3d7f7d681310ull, 0x17033d3c60c68173ull},
     {0x992d30ecfffffde5ull, 0x224698fc094cf91bull, 0x0000000000000000ull, 0x4000000000000000ull}}};
static const CurveSuite SUITE_VESTA = {
    "vesta",
    {0xc515ad7242eaa6b1ull, 0x9673928c7d01b212ull, 0x81639c4d96f78773ull, 0x267f9b2ee592271aull},
    {0xa70e2c1102b6d05full, 0x9bb97ea3c106f049ull, 0x9e5c4dfd492ae26eull, 0x2de6a9b8746d3f58ull},
    {{0x43cd42c800000001ull, 0x0205dd51cfa0961aull, 0x8e38e38e38e38e39ull, 0x38e38e38e38e38e3ull},
     {0x8b95c6aaf703bcc5ull, 0x216b8861ec72bd5dull, 0xacecf10f5f7c09a2ull, 0x1d935247b4473d17ull},
     {0xaeac67bbeb586a3dull, 0xd59d03d23b39cb11ull, 0xed7ee4a9cdf78f8full, 0x18760c7f7a9ad20dull},
     {0xfb539a6f0000002bull, 0xe1c521a795ac8356ull, 0x1c71c71c71c71c71ull, 0x31c71c71c71c71c7ull},
     {0xb7284f7eaf21a2e9ull, 0xa3ad678129b604d3ull, 0x1454798a5b5c56b2ull, 0x0a2de485568125d5ull},
     {0xf169c187d2533465ull, 0x30cd6d53df49d235ull, 0x0c621de8b91c242aull, 0x14735171ee542778ull},
     {0x6bef1642aaaaaaabull, 0x5601f4709a8adcb3ull, 0xda12f684bda12f68ull, 0x12f684bda12f684bull},
     {0x8bee58e5fb81de63ull, 0x21d910aefb03b31dull, 0xd6767887afbe04d1ull, 0x2ec9a923da239e8bull},
     {0x4986913ab4443034ull, 0x97a3ca5c24e9ea63ull, 0x66d1466e9de10e64ull, 0x19b0d87e16e25788ull},
     {0x8f64842c55555533ull, 0x8bc32d36fb21a6a3ull, 0x425ed097b425ed09ull, 0x1ed097b425ed097bull},
     {0x58dfecce86b2745eull, 0x06a767bfc35b5bacull, 0x9e7eb64f890a820cull, 0x2f44d6c801c1b8bfull},
     {0xd43d449776f99d2full, 0x926847fb9ddd76a1ull, 0x252659ba2b546c7eull, 0x3d59f455cafc7668ull},
     {0x8c46eb20fffffde5ull, 0x224698fc0994a8ddull, 0x0000000000000000ull, 0x4000000000000000ull}}};

template <class PR> inline void put_mont(uint32_t* dst, const uint64_t* canonical) {
  Fe<PR> m = fe_to_mont(fe_from_limbs64<PR>(canonical));
  for (int i = 0; i < 8; ++i) dst[i] = m.v[i];
}
template <class PR>
inline void fill_consts(const CurveSuite& S, H2cConsts& K) {
  put_mont<PR>(K.a, S.a);
  const uint64_t b[4] = {1265, 0, 0, 0}, thirteen[4] = {13, 0, 0, 0};
  put_mont<PR>(K.b, b);
  Fe<PR> z = fe_neg(fe_to_mont(fe_from_limbs64<PR>(thirteen)));
  for (int i = 0; i < 8; ++i) K.z[i] = z.v[i];
  for (int j = 0; j < 13; ++j) put_mont<PR>(K.iso[j], S.iso[j]);
  put_mont<PR>(K.root, S.root);
  Fe<PR> r2; for (int i = 0; i < 8; ++i) r2.v[i] = PR::r2(i);
  Fe<PR> r3 = fe_mul(r2, r2);
  for (int i = 0; i < 8; ++i) K.r3[i] = r3.v[i];
  // T = (p - 1) >> 32; (T - 1) / 2 = T >> 1 (T is odd)
  uint32_t t[8];
  for (int i = 0; i < 8; ++i) t[i] = i + 1 < 8 ? PR::p(i + 1) : 0;
  for (int i = 0; i < 8; ++i) K.t_minus1_over2[i] = (t[i] >> 1) | (i + 1 < 8 ? (t[i + 1] << 31) : 0);
}

// false when the domain prefix is too long
inline bool make_consts(bool pallas, const char* domain_prefix, H2cConsts& K) {
  memset(&K, 0, sizeof(K));
  const CurveSuite& S = pallas ? SUITE_PALLAS : SUITE_VESTA;
  if (pallas) fill_consts<FpParams>(S, K); else fill_consts<FqParams>(S, K);
  const size_t lp = strlen(domain_prefix), lc = strlen(S.id);
  if (lp >= 256 || 22 + lc + lp >= 256) return false;
  size_t o = 0;
  memcpy(K.dst_prime + o, domain_prefix, lp); o += lp;
  K.dst_prime[o++] = '-';
  memcpy(K.dst_prime + o, S.id, lc); o += lc;
  memcpy(K.dst_prime + o, "_XMD:BLAKE2b_SSWU_RO_", 21); o += 21;
  K.dst_prime[o++] = (uint8_t)(22 + lc + lp);
  K.dst_len = (uint32_t)o;
  return true;
}

}  // namespace h2c
