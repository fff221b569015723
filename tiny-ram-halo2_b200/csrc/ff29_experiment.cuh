// EXPERIMENT: reduced-radix (9 x 29-bit limbs, R' = 2^261) Montgomery multiplication with carry-free column
// accumulation: every MAC is a plain IMAD.WIDE.U32 (full rate); no IMAD.WIDE.X carry chains (measured half rate).
#pragma once
#include <stdint.h>
#if defined(__CUDACC__)
#define F29_HD __host__ __device__ __forceinline__
#else
#define F29_HD inline
#endif
namespace ff29 {
constexpr uint32_t MASK29 = (1u << 29) - 1;
struct FpP { static constexpr uint32_t P1 = 0x9698768u, P2 = 0x133e46e6u, P3 = 0xd31f812u, P4 = 0x224u; };
struct FqP { static constexpr uint32_t P1 = 0x2375908u, P2 = 0x52a3763u, P3 = 0xd31f813u, P4 = 0x224u; };
template <class PR> struct F29 { uint32_t l[9]; };

// r = a * b / 2^261 mod p (lazy: r < 2p, limbs < 2^29); inputs may have limbs up to 2^30
template <class PR> F29_HD F29<PR> mul(const F29<PR>& a, const F29<PR>& b) {
  uint64_t t[18];
#pragma unroll
  for (int k = 0; k < 18; ++k) t[k] = 0;
#pragma unroll
  for (int i = 0; i < 9; ++i)
#pragma unroll
    for (int j = 0; j < 9; ++j) t[i + j] += (uint64_t)a.l[i] * b.l[j];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    uint32_t m = (0u - (uint32_t)t[k]) & MASK29;
    t[k] += m;
    t[k + 1] += (uint64_t)m * PR::P1 + (t[k] >> 29);
    t[k + 2] += (uint64_t)m * PR::P2;
    t[k + 3] += (uint64_t)m * PR::P3;
    t[k + 4] += (uint64_t)m * PR::P4;
    t[k + 8] += (uint64_t)m << 22;
  }
  F29<PR> r;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    r.l[i] = (uint32_t)t[9 + i] & MASK29;
    t[10 + i] += t[9 + i] >> 29;
  }
  r.l[8] = (uint32_t)t[17];
  return r;
}

F29_HD uint64_t madw(uint32_t a, uint32_t b, uint64_t c) {
#ifdef __CUDA_ARCH__
  uint64_t d;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
  return d;
#else
  return (uint64_t)a * b + c;
#endif
}
F29_HD uint64_t mulw(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  uint64_t d;
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(d) : "r"(a), "r"(b));
  return d;
#else
  return (uint64_t)a * b;
#endif
}

template <class PR> F29_HD F29<PR> mul_v3(const F29<PR>& a, const F29<PR>& b) {
  uint64_t t[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) t[j] = mulw(a.l[0], b.l[j]);
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    uint32_t lo = (uint32_t)t[0];
    uint32_t m = (0u - lo) & MASK29;
    uint64_t c = (t[0] + m) >> 29;
    uint64_t n0 = madw(m, PR::P1, t[1]) + c;
    uint64_t n1 = madw(m, PR::P2, t[2]);
    uint64_t n2 = madw(m, PR::P3, t[3]);
    uint64_t n3 = madw(m, PR::P4, t[4]);
    uint64_t n7 = madw(m, 1u << 22, t[8]);
    t[0] = n0; t[1] = n1; t[2] = n2; t[3] = n3; t[4] = t[5]; t[5] = t[6]; t[6] = t[7]; t[7] = n7;
    if (i < 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] = madw(a.l[i + 1], b.l[j], t[j]);
      t[8] = mulw(a.l[i + 1], b.l[8]);
    }
  }
  F29<PR> r;
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    r.l[i] = (uint32_t)t[i] & MASK29;
    t[i + 1] += t[i] >> 29;
  }
  r.l[7] = (uint32_t)t[7] & MASK29;
  r.l[8] = (uint32_t)(t[7] >> 29);
  return r;
}

// row-wise (CIOS) variant: one row of a_i * b plus one reduction round per step, columns slide down by one limb.
// Column values stay < 2^61 so the inter-column carry fits 32 bits.
template <class PR> F29_HD F29<PR> mul_cios(const F29<PR>& a, const F29<PR>& b) {
  uint64_t t[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) t[j] = (uint64_t)a.l[0] * b.l[j];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    uint32_t m = (0u - (uint32_t)t[0]) & MASK29;
    // (t0 + m) has 29 zero low bits; carry = (t0 + m) >> 29
    uint32_t c = (uint32_t)((t[0] + m) >> 29);
    uint64_t n0 = t[1] + (uint64_t)m * PR::P1 + c;
    uint64_t n1 = t[2] + (uint64_t)m * PR::P2;
    uint64_t n2 = t[3] + (uint64_t)m * PR::P3;
    uint64_t n3 = t[4] + (uint64_t)m * PR::P4;
    uint64_t n7 = t[8] + (uint64_t)m * (1u << 22);
    t[0] = n0; t[1] = n1; t[2] = n2; t[3] = n3; t[4] = t[5]; t[5] = t[6]; t[6] = t[7]; t[7] = n7; t[8] = 0;
    if (i < 8) {
#pragma unroll
      for (int j = 0; j < 9; ++j) t[j] += (uint64_t)a.l[i + 1] * b.l[j];
    }
  }
  F29<PR> r;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    r.l[i] = (uint32_t)t[i] & MASK29;
    t[i + 1] += (uint32_t)(t[i] >> 29);
  }
  r.l[8] = (uint32_t)t[8];
  return r;
}
}  // namespace ff29
