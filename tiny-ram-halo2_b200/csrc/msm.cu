// K3: Pippenger multi-scalar multiplication on Pallas / Vesta for sm_100a.
//
// Replaces halo2_proofs::arithmetic::best_multiexp (and Params::commit / commit_lagrange which call it),
// halo2_proofs 0.2.0 @ a95945254dcc (Cargo.lock:619-621), reached from the reference at
// /root/reference/src/test_utils.rs:23,25,41.  The CPU routine runs one unsigned-window Pippenger per rayon
// thread slice; the group element it returns is unique, so only the algorithm's RESULT is shared with it.
//
// Pipeline per MSM (all on device, no host synchronisation):
//   1. hist      scalar -> canonical -> signed c-bit digits; warp-aggregated histogram of bucket sizes
//   2. scan      exclusive prefix sum -> bucket offsets
//   3. scatter   counting sort of (point index | sign) entries into bucket order
//   4. accumulate  multi-level segmented reduction: level 1 cuts the sorted entry list into aligned windows of 64 entries
//                (mixed XYZZ adds; one partial per bucket a window meets), the levels above add <= 16 partials of one bucket
//                (full adds), so run time is independent of how skewed the bucket sizes are (TinyRAM columns are mostly
//                0/1: one bucket gets ~n entries)
//   5. reduce    sum_b (b+1) * B_b per bucket set: chunked running sums, then the chunks' totals either scaled by their first
//                bucket index (c < 18) or folded by the two-level weighted sum of bucket_reduce.cuh (c >= 18)
//   6. final     Horner over bucket sets (none when bases are precomputed) -> Jacobian point
//
// Bases are static (Params.g_lagrange ++ [w]); with 180 GB of HBM the loader precomputes 2^(c*w) * P_i for
// every window w so that all windows share ONE bucket set: no per-window reduction, no doubling chain, long
// uniform bucket runs.  When that table would not fit the budget the windows keep separate bucket sets.
#include "common.cuh"
#include "ec.cuh"
#include "bucket_reduce.cuh"
#include "scan.cuh"

#include <algorithm>
#include <cstdlib>

using namespace ff;
using namespace ec;
using scan::run_scan;
using scan::SCAN_BLOCK;

struct MsmGeom {
  unsigned c;        // window bits
  unsigned W;        // number of windows
  unsigned nsets;    // bucket sets: 1 (precomputed bases) or W
  unsigned B;        // buckets per set = 2^(c-1)
  unsigned nb;       // nsets * B
  unsigned precomp;  // 1: entry index = w * stride + i into the precomputed table
  size_t stride;     // number of loaded bases (table row length)
};

struct trp_bases_impl {
  trp_bases pub;
  MsmGeom g;
};

namespace {

constexpr unsigned L1 = 64;     // entries per level-1 task
constexpr unsigned L2 = 16;     // partials per task at levels >= 2
constexpr unsigned RED_S = 16;  // buckets per running-sum chunk
constexpr int ACC_THREADS = 128;

__device__ __forceinline__ unsigned extract_bits(const uint32_t* v, unsigned off, unsigned c) {
  unsigned idx = off >> 5, sh = off & 31;
  if (idx >= 8) return 0;
  uint64_t lo = v[idx];
  uint64_t hi = idx + 1 < 8 ? v[idx + 1] : 0;
  uint64_t x = (lo | (hi << 32)) >> sh;
  return (unsigned)(x & ((1u << c) - 1));
}

// Signed-digit recoding shared by hist and scatter: digit d_w in [-2^(c-1), 2^(c-1)), sum d_w 2^(cw) = scalar.
// f(w, key, neg) is called for every window by every lane (key = 0xffffffff for zero digits) so that
// warp-collective code inside f stays converged.
template <class SPR, class F>
__device__ __forceinline__ void for_each_digit(const uint4* scalars, size_t i, size_t n, const MsmGeom& g, F f) {
  uint32_t v[8];
  if (i < n) {
    Fe<SPR> s = fe_from_mont(fe_load<SPR>(scalars + 2 * i));
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = s.v[k];
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = 0;
  }
  // a warp whose 32 scalars are all zero has nothing to sort (TinyRAM advice columns are zero outside the 2^16-row tables:
  // 94 % of the warps at k = 20); the exit is warp-uniform, so the collectives inside f stay converged
  if (!__any_sync(0xffffffffu, (v[0] | v[1] | v[2] | v[3] | v[4] | v[5] | v[6] | v[7]) != 0)) return;
  unsigned carry = 0;
  const unsigned half = 1u << (g.c - 1);
  for (unsigned w = 0; w < g.W; ++w) {
    unsigned raw = extract_bits(v, w * g.c, g.c) + carry;
    unsigned neg = raw >= half;
    unsigned mag = neg ? (2 * half - raw) : raw;
    carry = neg;
    unsigned key = mag ? ((g.precomp ? 0u : w * g.B) + mag - 1) : 0xffffffffu;
    f(w, key, neg);
  }
}

// Batches are processed column-concurrently: blockIdx.y is the column, its buckets are [col * nb, (col + 1) * nb).
template <class SPR>
__global__ void msm_hist_kernel(const uint4* scalars, size_t n, MsmGeom g, uint32_t* counts) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  scalars += 2 * (size_t)blockIdx.y * n;
  counts += (size_t)blockIdx.y * g.nb;
  for_each_digit<SPR>(scalars, i, n, g, [&](unsigned, unsigned key, unsigned) {
    unsigned peers = __match_any_sync(0xffffffffu, key);
    if (key != 0xffffffffu && (unsigned)(__ffs(peers) - 1) == (threadIdx.x & 31)) atomicAdd(&counts[key], __popc(peers));
  });
}

template <class SPR>
__global__ void msm_scatter_kernel(const uint4* scalars, size_t n, MsmGeom g, uint32_t* cursor, uint32_t* entries) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31;
  scalars += 2 * (size_t)blockIdx.y * n;
  cursor += (size_t)blockIdx.y * g.nb;
  for_each_digit<SPR>(scalars, i, n, g, [&](unsigned w, unsigned key, unsigned neg) {
    unsigned peers = __match_any_sync(0xffffffffu, key);
    unsigned leader = __ffs(peers) - 1;
    unsigned base = 0;
    if (key != 0xffffffffu && leader == lane) base = atomicAdd(&cursor[key], __popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (key != 0xffffffffu) {
      unsigned rank = __popc(peers & ((1u << lane) - 1));
      uint32_t idx = (uint32_t)(g.precomp ? (size_t)w * g.stride + i : i);
      entries[base + rank] = idx | (neg << 31);
    }
  });
}

// ---- accumulation ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned find_segment(const uint32_t* off, unsigned nseg, uint32_t t) {
  // largest b in [0, nseg) with off[b] <= t   (off is non-decreasing, off[nseg] > t)
  unsigned lo = 0, hi = nseg;
  while (hi - lo > 1) {
    unsigned mid = (lo + hi) >> 1;
    if (__ldg(off + mid) <= t) lo = mid; else hi = mid;
  }
  return lo;
}

template <class BPR>
__device__ __forceinline__ Affine<BPR> load_affine(const uint4* bases, size_t idx) {
  const uint4* p = bases + 4 * idx;
  uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
  Affine<BPR> r;
  r.x.v[0] = a.x; r.x.v[1] = a.y; r.x.v[2] = a.z; r.x.v[3] = a.w; r.x.v[4] = b.x; r.x.v[5] = b.y; r.x.v[6] = b.z; r.x.v[7] = b.w;
  r.y.v[0] = c.x; r.y.v[1] = c.y; r.y.v[2] = c.z; r.y.v[3] = c.w; r.y.v[4] = d.x; r.y.v[5] = d.y; r.y.v[6] = d.z; r.y.v[7] = d.w;
  return r;
}
template <class BPR>
__device__ __forceinline__ XYZZ<BPR> load_xyzz(const uint4* p) {
  XYZZ<BPR> r;
  r.x = fe_load<BPR>(p); r.y = fe_load<BPR>(p + 2); r.zz = fe_load<BPR>(p + 4); r.zzz = fe_load<BPR>(p + 6);
  return r;
}
template <class BPR>
__device__ __forceinline__ void store_xyzz(uint4* p, const XYZZ<BPR>& a) {
  fe_store(p, a.x); fe_store(p + 2, a.y); fe_store(p + 4, a.zz); fe_store(p + 6, a.zzz);
}

// Level 1: entries (point index | sign) -> affine bases, mixed adds.  Tasks cut the SORTED ENTRY LIST into aligned windows of L1
// entries (not every bucket into tasks of its own, round 1's layout, whose last task per bucket was partial: ~6 % idle lanes at
// 512 entries per bucket, ~30 % at the 26 entries per bucket of c = 20), so every thread of a warp performs exactly L1
// additions whatever the bucket sizes are.  A thread emits one partial per bucket its window meets:
// bucket b's partials are the slots pbase[b] + (t - off[b] / L1), pbase = the scan of scan_input mode 2, which is exactly the
// per-bucket layout the upper levels read.  Measured on B200 (profiles/msm_variants_r02.md): 8 x 2^20 uniform, c = 16:
// accumulate 22.2 -> 21.3 ms; it is what makes wider windows pay (c = 20: 25.0 -> 18.0 ms).
template <class BPR>
__global__ void __launch_bounds__(ACC_THREADS) msm_accum_l1_seg_kernel(const uint32_t* entries, const uint32_t* off, const uint32_t* pbase,
                                                                     unsigned nb, const uint4* bases, uint4* partials) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t total = __ldg(off + nb);
  if ((uint64_t)t * L1 >= total) return;
  const uint32_t e0 = t * L1, e1 = (uint32_t)min((uint64_t)e0 + L1, (uint64_t)total);
  unsigned b = find_segment(off, nb, e0);              // the (non-empty) bucket holding entry e0
  uint32_t next = __ldg(off + b + 1);
  XYZZ<BPR> acc = xyzz_identity<BPR>();
  for (uint32_t e = e0; e < e1; ++e) {
    if (e == next) {                                   // bucket boundary inside the window: emit, move to the next non-empty bucket
      store_xyzz(partials + 8 * (size_t)(__ldg(pbase + b) + (t - __ldg(off + b) / L1)), acc);
      acc = xyzz_identity<BPR>();
      do { ++b; next = __ldg(off + b + 1); } while (next <= e);
    }
    uint32_t ent = __ldg(entries + e);
    Affine<BPR> p = load_affine<BPR>(bases, ent & 0x7fffffffu);
    if (ent >> 31) p.y = fe_neg(p.y);
    xyzz_add_mixed(acc, p);
  }
  store_xyzz(partials + 8 * (size_t)(__ldg(pbase + b) + (t - __ldg(off + b) / L1)), acc);
}

// levels >= 2: partial sums of the previous level, full adds
template <class BPR>
__global__ void __launch_bounds__(ACC_THREADS) msm_accum_ln_kernel(const uint4* in, const uint32_t* in_off,
                                                                 const uint32_t* task_off, unsigned nb, uint4* partials) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= __ldg(task_off + nb)) return;
  unsigned b = find_segment(task_off, nb, t);
  uint32_t j = t - __ldg(task_off + b);
  uint32_t start = __ldg(in_off + b) + j * L2;
  uint32_t end = min(start + L2, __ldg(in_off + b + 1));
  XYZZ<BPR> acc = load_xyzz<BPR>(in + 8 * (size_t)start);
  for (uint32_t e = start + 1; e < end; ++e) xyzz_add(acc, load_xyzz<BPR>(in + 8 * (size_t)e));
  store_xyzz(partials + 8 * (size_t)t, acc);
}

// ---- bucket reduction -------------------------------------------------------------------------------------
// After the last level every bucket has 0 or 1 partial: B_b = part[off[b]] if off[b+1] > off[b] else identity.
template <class BPR>
__device__ __forceinline__ XYZZ<BPR> bucket_value(const uint4* part, const uint32_t* off, unsigned b) {
  uint32_t o = __ldg(off + b);
  if (__ldg(off + b + 1) > o) return load_xyzz<BPR>(part + 8 * (size_t)o);
  return xyzz_identity<BPR>();
}

template <class BPR>
__device__ __forceinline__ XYZZ<BPR> xyzz_mul_small(const XYZZ<BPR>& p, unsigned k) {
  XYZZ<BPR> acc = xyzz_identity<BPR>();
  if (k == 0 || xyzz_is_identity(p)) return acc;
  int top = 31 - __clz(k);
  for (int bit = top; bit >= 0; --bit) {
    xyzz_dbl(acc);
    if ((k >> bit) & 1) xyzz_add(acc, p);
  }
  return acc;
}

// chunk of RED_S consecutive buckets [lo, lo+S) of one set: sum_{b} (b+1) B_b = running-sum part + lo * (sum B_b)
template <class BPR>
__global__ void __launch_bounds__(ACC_THREADS) msm_reduce_chunks_kernel(const uint4* part, const uint32_t* off, MsmGeom g, unsigned total_sets, uint4* chunk_out) {
  unsigned chunks_per_set = g.B / RED_S ? g.B / RED_S : 1;
  unsigned S = g.B < RED_S ? g.B : RED_S;
  unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total_sets * chunks_per_set) return;
  unsigned set = t / chunks_per_set, ch = t % chunks_per_set;
  unsigned lo = ch * S;
  XYZZ<BPR> run = xyzz_identity<BPR>(), acc = xyzz_identity<BPR>();
  for (int k = (int)S - 1; k >= 0; --k) {
    XYZZ<BPR> bv = bucket_value<BPR>(part, off, set * g.B + lo + k);
    xyzz_add(run, bv);
    xyzz_add(acc, run);
  }
  XYZZ<BPR> scaled = xyzz_mul_small(run, lo);
  xyzz_add(acc, scaled);
  store_xyzz(chunk_out + 8 * (size_t)t, acc);
}

// one block per set: sum the chunk results -> set_sums[set]
template <class BPR>
__global__ void __launch_bounds__(256) msm_reduce_sets_kernel(const uint4* chunk_in, unsigned chunks_per_set, uint4* set_sums) {
  __shared__ uint4 sm[256 * 8];
  unsigned set = blockIdx.x;
  const uint4* in = chunk_in + 8 * (size_t)set * chunks_per_set;
  XYZZ<BPR> acc = xyzz_identity<BPR>();
  for (unsigned k = threadIdx.x; k < chunks_per_set; k += blockDim.x) xyzz_add(acc, load_xyzz<BPR>(in + 8 * (size_t)k));
  store_xyzz(sm + 8 * threadIdx.x, acc);
  __syncthreads();
  for (unsigned s = blockDim.x >> 1; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      XYZZ<BPR> a = load_xyzz<BPR>(sm + 8 * threadIdx.x);
      xyzz_add(a, load_xyzz<BPR>(sm + 8 * (threadIdx.x + s)));
      store_xyzz(sm + 8 * threadIdx.x, a);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    uint4* o = set_sums + 8 * (size_t)set;
    for (int k = 0; k < 8; ++k) o[k] = sm[k];
  }
}

// Two-level reduction (the arithmetic is bucket_reduce.cuh, host-tested; default for c >= 18, TRP_MSM_REDUCE=1/2 forces one
// or the other): chunks emit their own weighted sum AND their total; the per-set kernel turns the totals into sum_ch ch * tot_ch
// by running sums inside a thread's run of chunks and a (sum, weighted sum) tree across threads, instead of one doubling chain
// per chunk.  Measured (profiles/msm_variants_r02.md, 8 x 2^20): c = 16 0.63 -> 1.22 ms (worse: few chunks per set), c = 18
// 1.86 -> 1.77, c = 20 8.33 -> 3.95 ms.
constexpr unsigned LOG_RED_S = 4;            // RED_S = 16
constexpr int RED2_THREADS = 128;
static_assert((1u << LOG_RED_S) == RED_S, "LOG_RED_S");
template <class BPR>
__global__ void __launch_bounds__(ACC_THREADS) msm_reduce_chunks2_kernel(const uint4* part, const uint32_t* off, MsmGeom g, unsigned total_sets,
                                                                       uint4* chunk_acc, uint4* chunk_tot) {
  const unsigned chunks_per_set = g.B / RED_S;
  unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total_sets * chunks_per_set) return;
  unsigned set = t / chunks_per_set, lo = (t % chunks_per_set) * RED_S;
  XYZZ<BPR> run = xyzz_identity<BPR>(), acc = xyzz_identity<BPR>();
  for (int k = (int)RED_S - 1; k >= 0; --k) {
    xyzz_add(run, bucket_value<BPR>(part, off, set * g.B + lo + k));
    xyzz_add(acc, run);
  }
  store_xyzz(chunk_acc + 8 * (size_t)t, acc);
  store_xyzz(chunk_tot + 8 * (size_t)t, run);
}

// block b sums the RED2_THREADS << log_m chunks that start at chunk b * (RED2_THREADS << log_m) into one node (G blocks per set)
template <class BPR> __device__ __forceinline__ void put_node(uint4* p, const WNode<BPR>& n) {
  store_xyzz(p, n.a); store_xyzz(p + 8, n.s); store_xyzz(p + 16, n.w);
}
template <class BPR> __device__ __forceinline__ WNode<BPR> get_node(const uint4* p) {
  WNode<BPR> n;
  n.a = load_xyzz<BPR>(p); n.s = load_xyzz<BPR>(p + 8); n.w = load_xyzz<BPR>(p + 16);
  return n;
}
template <class BPR>
__global__ void __launch_bounds__(RED2_THREADS) msm_reduce_sets2_kernel(const uint4* chunk_acc, const uint4* chunk_tot, unsigned log_m, uint4* nodes) {
  __shared__ uint4 sm[RED2_THREADS * 24];    // a, s, w per thread
  const unsigned m = 1u << log_m, tid = threadIdx.x;
  const size_t first = ((size_t)blockIdx.x * RED2_THREADS + tid) << log_m;
  const uint4* acc = chunk_acc + 8 * first;
  const uint4* tot = chunk_tot + 8 * first;
  put_node<BPR>(sm + 24 * tid, wnode_leaf<BPR>(m, [&](unsigned ch) { return load_xyzz<BPR>(acc + 8 * (size_t)ch); },
                                               [&](unsigned ch) { return load_xyzz<BPR>(tot + 8 * (size_t)ch); }));
  __syncthreads();
  unsigned lv = 0;
  for (unsigned stride = 1; stride < RED2_THREADS; stride <<= 1, ++lv) {
    if ((tid & (2 * stride - 1)) == 0) {
      WNode<BPR> l = get_node<BPR>(sm + 24 * tid);
      wnode_combine(l, get_node<BPR>(sm + 24 * (tid + stride)), log_m + lv);
      put_node<BPR>(sm + 24 * tid, l);
    }
    __syncthreads();
  }
  if (tid < 24) nodes[24 * (size_t)blockIdx.x + tid] = sm[tid];
}

// one block per set: the G = 2^log_g nodes of the set (each covering 2^log_len chunks) -> set_sums[set]
template <class BPR>
__global__ void __launch_bounds__(RED2_THREADS) msm_reduce_nodes_kernel(uint4* nodes, unsigned log_g, unsigned log_len, uint4* set_sums) {
  const unsigned G = 1u << log_g, tid = threadIdx.x;
  uint4* mine = nodes + 24 * ((size_t)blockIdx.x << log_g);
  unsigned lv = 0;
  for (unsigned stride = 1; stride < G; stride <<= 1, ++lv) {
    if (tid < G && (tid & (2 * stride - 1)) == 0) {
      WNode<BPR> l = get_node<BPR>(mine + 24 * tid);
      wnode_combine(l, get_node<BPR>(mine + 24 * (tid + stride)), log_len + lv);
      put_node<BPR>(mine + 24 * tid, l);
    }
    __syncthreads();     // one block per set: its threads are the only readers and writers of the set's nodes
  }
  if (tid == 0) store_xyzz(set_sums + 8 * (size_t)blockIdx.x, wnode_root(get_node<BPR>(mine), LOG_RED_S));
}

// Horner over sets (window w has weight 2^(c*w)) and conversion XYZZ -> Jacobian (X*ZZ^4... no inversion):
// (X', Y', Z') = (X * ZZ^4, Y * ZZZ^4, ZZ * ZZZ) since Z'^2 = ZZ^5 and Z'^3 = ZZZ^5.
template <class BPR>
__global__ void msm_final_kernel(const uint4* set_sums, MsmGeom g, unsigned mcols, uint4* out_jac) {
  unsigned col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= mcols) return;
  set_sums += 8 * (size_t)col * g.nsets;
  out_jac += 6 * (size_t)col;
  XYZZ<BPR> acc = load_xyzz<BPR>(set_sums + 8 * (size_t)(g.nsets - 1));
  for (int w = (int)g.nsets - 2; w >= 0; --w) {
    for (unsigned k = 0; k < g.c; ++k) xyzz_dbl(acc);
    xyzz_add(acc, load_xyzz<BPR>(set_sums + 8 * (size_t)w));
  }
  Fe<BPR> X, Y, Z;
  if (xyzz_is_identity(acc)) {
    X = fe_zero<BPR>(); Y = fe_zero<BPR>(); Z = fe_zero<BPR>();
  } else {
    Fe<BPR> zz2 = fe_sqr(acc.zz), zzz2 = fe_sqr(acc.zzz);
    X = fe_mul(acc.x, fe_sqr(zz2));
    Y = fe_mul(acc.y, fe_sqr(zzz2));
    Z = fe_mul(acc.zz, acc.zzz);
  }
  fe_store(out_jac, X); fe_store(out_jac + 2, Y); fe_store(out_jac + 4, Z);
}

// batch normalisation of the m results: (X, Y, Z) -> (X/Z^2, Y/Z^3, 1), identity -> (0, 0, 0).  Makes the output
// canonical (the accumulation order inside buckets is not deterministic, the group element is).
template <class BPR>
__global__ void msm_normalize_kernel(uint4* out_jac, size_t m) {
  size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  uint4* o = out_jac + 6 * k;
  Fe<BPR> X = fe_load<BPR>(o), Y = fe_load<BPR>(o + 2), Z = fe_load<BPR>(o + 4);
  if (fe_is_zero(Z)) {
    for (int q = 0; q < 6; ++q) o[q] = make_uint4(0, 0, 0, 0);
    return;
  }
  Fe<BPR> zi = fe_inv(Z), zi2 = fe_sqr(zi);
  fe_store(o, fe_mul(X, zi2));
  fe_store(o + 2, fe_mul(Y, fe_mul(zi2, zi)));
  fe_store(o + 4, fe_one<BPR>());
}

// ---- base precomputation: table[w][i] = 2^(c*w) * P_i, affine ---------------------------------------------------
// One thread per base: walk the doubling chain once (kept in local memory), then normalise all W-1 multiples
// with a single inversion (Montgomery's trick); the prefix products are stashed in the output slots.
constexpr unsigned PRECOMP_MAX_W = 64;
template <class BPR>
__global__ void __launch_bounds__(128) msm_precompute_kernel(uint4* table, size_t n, unsigned c, unsigned W) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<BPR> p = load_affine<BPR>(table, i);
  if (affine_is_identity(p)) {
    for (unsigned w = 1; w < W; ++w)
      for (int k = 0; k < 4; ++k) table[4 * (w * n + i) + k] = make_uint4(0, 0, 0, 0);
    return;
  }
  XYZZ<BPR> chain[PRECOMP_MAX_W - 1];
  XYZZ<BPR> q = xyzz_from_affine(p);
  Fe<BPR> prod = fe_one<BPR>();
  for (unsigned w = 1; w < W; ++w) {
    for (unsigned k = 0; k < c; ++k) xyzz_dbl(q);     // odd prime order: never reaches the identity
    chain[w - 1] = q;
    fe_store(table + 4 * (w * n + i), prod);          // prefix product of the elements before w
    prod = fe_mul(prod, fe_mul(q.zz, q.zzz));
  }
  Fe<BPR> inv = fe_inv(prod);
  for (unsigned w = W - 1; w >= 1; --w) {
    uint4* slot = table + 4 * (w * n + i);
    const XYZZ<BPR> e = chain[w - 1];
    Fe<BPR> einv = fe_mul(inv, fe_load<BPR>(slot));   // 1 / (zz * zzz) of element w
    inv = fe_mul(inv, fe_mul(e.zz, e.zzz));
    fe_store(slot, fe_mul(e.x, fe_mul(einv, e.zzz)));       // X / ZZ
    fe_store(slot + 2, fe_mul(e.y, fe_mul(einv, e.zz)));    // Y / ZZZ
  }
}

// ---- synthetic input generator: out[i] = P0 + i * D (affine), used by bench.py / tests to make bases on device -------
constexpr unsigned GEN_CH = 16;
template <class BPR>
__global__ void __launch_bounds__(128) points_progression_kernel(Affine<BPR> p0, Affine<BPR> d, size_t n, uint4* out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t start = t * GEN_CH;
  if (start >= n) return;
  XYZZ<BPR> cur = xyzz_mul_small(xyzz_from_affine(d), (unsigned)start);
  xyzz_add_mixed(cur, p0);
  XYZZ<BPR> chain[GEN_CH];
  Fe<BPR> pre[GEN_CH];
  Fe<BPR> prod = fe_one<BPR>();
  unsigned cnt = (unsigned)(n - start < GEN_CH ? n - start : GEN_CH);
  for (unsigned k = 0; k < cnt; ++k) {
    chain[k] = cur;
    pre[k] = prod;
    if (!xyzz_is_identity(cur)) prod = fe_mul(prod, fe_mul(cur.zz, cur.zzz));
    xyzz_add_mixed(cur, d);
  }
  Fe<BPR> inv = fe_inv(prod);
  for (int k = (int)cnt - 1; k >= 0; --k) {
    uint4* slot = out + 4 * (start + k);
    const XYZZ<BPR> e = chain[k];
    if (xyzz_is_identity(e)) { for (int q = 0; q < 4; ++q) slot[q] = make_uint4(0, 0, 0, 0); continue; }
    Fe<BPR> einv = fe_mul(inv, pre[k]);
    inv = fe_mul(inv, fe_mul(e.zz, e.zzz));
    fe_store(slot, fe_mul(e.x, fe_mul(einv, e.zzz)));
    fe_store(slot + 2, fe_mul(e.y, fe_mul(einv, e.zz)));
  }
}

// Window width.  Measured on B200 with the segmented level 1 (profiles/msm_variants_r02.md, 8 columns per batch):
//   n = 2^20 uniform scalars:  c = 16 26.2 ms, c = 17 24.8 ms (best), c = 18 26.0, c = 20 28.1
//   n = 2^22 uniform (4 columns): c = 16 51.0 ms, c = 18 47.5, c = 20 44.9 (best measured)
//   n = 2^20, columns that are zero outside 2^16 rows (TinyRAM advice): c = 13 2.17 ms (best), c = 16 2.7, c = 17 3.3 -- the
//   bucket reduction's 2^(c-1) buckets per column are all that is left there, so callers that know their columns are sparse
//   ask for a narrow table (flags bits 8..15 of trp_bases_load_ex / trp_dev_bases_load_ex).
unsigned choose_c(size_t n, unsigned requested = 0) {
  unsigned lg = 0;
  while (((size_t)2 << lg) <= n) ++lg;   // floor(log2 n) for n >= 1
  int c = (int)lg - 4;
  if (c < 4) c = 4;
  if (c > 16) c = 16;
  if (lg >= 20) c = 17 + (int)(3 * (lg - 20)) / 2;      // 2^20: 17, 2^21: 18, 2^22: 20, 2^23: 21, 2^24: 23 -> capped
  if (c > 22) c = 22;
  if (requested >= 2 && requested <= 22) c = (int)requested;
  else if (const char* e = getenv("TRP_MSM_C")) { int v = atoi(e); if (v >= 2 && v <= 22) c = v; }     // the A/B sweep
  return (unsigned)c;
}

// Worst-case task counts per level for mc columns processed together (entries may all fall into one bucket, or
// spread over all of them).
std::vector<size_t> plan_levels(const MsmGeom& g, size_t n, size_t mc) {
  const size_t M = n * g.W * mc, NB = (size_t)g.nb * mc;
  std::vector<size_t> level_tasks;
  size_t single = (n * g.W + L1 - 1) / L1;      // tasks if ONE bucket of a column held all its entries
  ++single;                                     // an unaligned run of entries meets one window more
  size_t bound = (M + L1 - 1) / L1 + NB;        // sum_b ceil(cnt_b / L1) <= M / L1 + NB
  level_tasks.push_back(bound);
  while (single > 1) {
    single = (single + L2 - 1) / L2;
    bound = (bound + L2 - 1) / L2 + NB;
    level_tasks.push_back(bound);
  }
  return level_tasks;
}

struct MsmWs {
  uint32_t *counts, *offsets, *cursor, *block_sums, *total, *entries, *task_off[2];
  uint4 *part[2], *chunk_out, *nodes, *set_sums;
  size_t bytes;
};
MsmWs carve(const MsmGeom& g, size_t n, size_t mc, char* base) {
  const size_t M = n * g.W * mc, NB = (size_t)g.nb * mc;
  std::vector<size_t> lt = plan_levels(g, n, mc);
  unsigned chunks_per_set = g.B / RED_S ? g.B / RED_S : 1;
  WsCursor cur{base, 0, 0};
  MsmWs w;
  w.counts = cur.take<uint32_t>(NB + 1);
  w.offsets = cur.take<uint32_t>(NB + 1);
  w.cursor = cur.take<uint32_t>(NB + 1);
  w.block_sums = cur.take<uint32_t>(SCAN_BLOCK);
  w.total = cur.take<uint32_t>(4);
  w.entries = cur.take<uint32_t>(M ? M : 1);
  w.task_off[0] = cur.take<uint32_t>(NB + 1);
  w.task_off[1] = cur.take<uint32_t>(NB + 1);
  w.part[0] = cur.take<uint4>(8 * lt[0]);
  w.part[1] = cur.take<uint4>(8 * (lt.size() > 1 ? lt[1] : 1));
  w.chunk_out = cur.take<uint4>(2 * 8 * (size_t)g.nsets * mc * chunks_per_set);   // second half: the chunk totals of TRP_MSM_REDUCE=2
  w.nodes = cur.take<uint4>(24 * ((size_t)g.nsets * mc * chunks_per_set / RED2_THREADS + 1));   // and its per-block nodes
  w.set_sums = cur.take<uint4>(8 * (size_t)g.nsets * mc);
  w.bytes = cur.off + 4096;
  return w;
}

// mc columns processed together (all on device, no host synchronisation)
template <class BPR, class SPR>
int msm_chunk(trp_ctx* ctx, const trp_bases_impl* bs, const uint4* d_scalars, size_t n, size_t mc, uint4* d_out, char* ws, size_t ws_cap) {
  const MsmGeom& g = bs->g;
  const size_t M = n * g.W * mc, NB = (size_t)g.nb * mc;
  if (M >= ((size_t)1 << 32) - 1 || NB >= ((size_t)1 << 31)) TRP_FAIL(ctx, TRP_E_INVALID, "internal: MSM chunk of %zu columns is too large", mc);
  std::vector<size_t> level_tasks = plan_levels(g, n, mc);
  MsmWs w = carve(g, n, mc, ws);
  if (w.bytes > ws_cap) TRP_FAIL(ctx, TRP_E_INVALID, "internal: MSM workspace underestimated (%zu > %zu)", w.bytes, ws_cap);
  {
    ProfScope ps(ctx, PROF_MSM_SORT);
    TRP_CUDA(ctx, cudaMemsetAsync(w.counts, 0, (NB + 1) * sizeof(uint32_t), ctx->stream));
    dim3 sgrid((unsigned)((n + 127) / 128), (unsigned)mc);
    msm_hist_kernel<SPR><<<sgrid, 128, 0, ctx->stream>>>(d_scalars, n, g, w.counts);
    TRP_LAUNCHED(ctx);
    TRP_TRY(run_scan(ctx, w.counts, w.offsets, w.cursor, w.block_sums, w.total, NB, 0, 1));
    msm_scatter_kernel<SPR><<<sgrid, 128, 0, ctx->stream>>>(d_scalars, n, g, w.cursor, w.entries);
    TRP_LAUNCHED(ctx);
    TRP_TRY(run_scan(ctx, w.offsets, w.task_off[0], nullptr, w.block_sums, w.total, NB, 2, L1));
  }
  {
    ProfScope ps(ctx, PROF_MSM_ACCUM_L1, (double)M);      // work = entries if no digit were zero (n * windows * columns): an upper bound
    msm_accum_l1_seg_kernel<BPR><<<(unsigned)(((M + L1 - 1) / L1 + ACC_THREADS - 1) / ACC_THREADS), ACC_THREADS, 0, ctx->stream>>>(
        w.entries, w.offsets, w.task_off[0], (unsigned)NB, (const uint4*)bs->pub.d_xy, w.part[0]);
    TRP_LAUNCHED(ctx);
  }
  int curp = 0;
  {
    ProfScope ps(ctx, PROF_MSM_LEVELS);
    for (size_t lv = 1; lv < level_tasks.size(); ++lv) {
      TRP_TRY(run_scan(ctx, w.task_off[curp], w.task_off[curp ^ 1], nullptr, w.block_sums, w.total, NB, 1, L2));
      msm_accum_ln_kernel<BPR><<<(unsigned)((level_tasks[lv] + ACC_THREADS - 1) / ACC_THREADS), ACC_THREADS, 0, ctx->stream>>>(
          w.part[curp], w.task_off[curp], w.task_off[curp ^ 1], (unsigned)NB, w.part[curp ^ 1]);
      TRP_LAUNCHED(ctx);
      curp ^= 1;
    }
  }
  {
    ProfScope ps(ctx, PROF_MSM_REDUCE);
    unsigned chunks_per_set = g.B / RED_S ? g.B / RED_S : 1;
    unsigned total_sets = (unsigned)(g.nsets * mc);
    unsigned nchunks = total_sets * chunks_per_set;
    static const int reduce_env = [] { const char* e = getenv("TRP_MSM_REDUCE"); return e ? atoi(e) : 0; }();
    const bool two_level = reduce_env == 2 || (reduce_env != 1 && g.c >= 18);
    if (two_level && g.B >= RED_S * (unsigned)RED2_THREADS) {
      // chunks per set = RED2_THREADS * 2^log_m * 2^log_g: at most 4 chunks per thread, the rest as blocks (B is a power of two)
      unsigned log_cps = 0;
      while ((2u << log_cps) <= chunks_per_set) ++log_cps;
      const unsigned log_t = 7;                                                       // RED2_THREADS = 128
      unsigned log_m = log_cps - log_t < 2 ? log_cps - log_t : 2, log_g = log_cps - log_t - log_m;
      if (log_g > log_t) { log_m += log_g - log_t; log_g = log_t; }                  // at most RED2_THREADS nodes per set
      uint4* chunk_tot = w.chunk_out + 8 * (size_t)nchunks;
      msm_reduce_chunks2_kernel<BPR><<<(nchunks + ACC_THREADS - 1) / ACC_THREADS, ACC_THREADS, 0, ctx->stream>>>(w.part[curp], w.task_off[curp], g, total_sets, w.chunk_out, chunk_tot);
      TRP_LAUNCHED(ctx);
      msm_reduce_sets2_kernel<BPR><<<total_sets << log_g, RED2_THREADS, 0, ctx->stream>>>(w.chunk_out, chunk_tot, log_m, w.nodes);
      TRP_LAUNCHED(ctx);
      msm_reduce_nodes_kernel<BPR><<<total_sets, RED2_THREADS, 0, ctx->stream>>>(w.nodes, log_g, log_t + log_m, w.set_sums);
      TRP_LAUNCHED(ctx);
    } else {
      msm_reduce_chunks_kernel<BPR><<<(nchunks + ACC_THREADS - 1) / ACC_THREADS, ACC_THREADS, 0, ctx->stream>>>(w.part[curp], w.task_off[curp], g, total_sets, w.chunk_out);
      TRP_LAUNCHED(ctx);
      msm_reduce_sets_kernel<BPR><<<total_sets, 256, 0, ctx->stream>>>(w.chunk_out, chunks_per_set, w.set_sums);
      TRP_LAUNCHED(ctx);
    }
    msm_final_kernel<BPR><<<(unsigned)((mc + 31) / 32), 32, 0, ctx->stream>>>(w.set_sums, g, (unsigned)mc, d_out);
    TRP_LAUNCHED(ctx);
  }
  return TRP_OK;
}

// columns processed together: bounded by the 32-bit entry index and by a scratch budget
size_t msm_cols_per_chunk(const MsmGeom& g, size_t n, size_t m) {
  size_t budget = (size_t)6 << 30;
  if (const char* e = getenv("TRP_MSM_BATCH_MB")) budget = (size_t)atoll(e) << 20;
  size_t mc = m ? m : 1;
  if (n) {
    size_t cap = (((size_t)1 << 32) - 2) / (n * g.W);
    if (cap < 1) cap = 1;
    if (mc > cap) mc = cap;
  }
  while (mc > 1 && (carve(g, n, mc, nullptr).bytes > budget || (size_t)g.nb * mc > (size_t)SCAN_BLOCK * SCAN_BLOCK)) mc = (mc + 1) / 2;
  return mc;
}

size_t msm_ws_bytes(const MsmGeom& g, size_t n, size_t m) {
  return carve(g, n, msm_cols_per_chunk(g, n, m), nullptr).bytes;
}

}  // namespace

// ---- C-ABI facing implementation -------------------------------------------------------------------------------
int trp_bases_create(trp_ctx* ctx, const void* src, bool src_on_device, size_t n, int flags, trp_bases** out) {
  if (!out) TRP_FAIL(ctx, TRP_E_INVALID, "out is NULL");
  if (n > ((size_t)1 << 27)) TRP_FAIL(ctx, TRP_E_INVALID, "too many bases (%zu)", n);
  trp_bases_impl* b = new trp_bases_impl();
  b->pub.ctx = ctx; b->pub.n = n; b->pub.d_xy = nullptr;
  MsmGeom& g = b->g;
  g.c = choose_c(n ? n : 1, ((unsigned)flags >> 8) & 0xff);
  g.W = (256 + g.c - 1) / g.c;
  g.B = 1u << (g.c - 1);
  g.stride = n;
  size_t budget = (size_t)48 << 30;
  if (const char* e = getenv("TRP_MSM_PRECOMP_BUDGET_MB")) budget = (size_t)atoll(e) << 20;
  bool precomp = (flags & 1) ? false : ((flags & 2) ? true : (n * 64 * g.W <= budget));
  if ((size_t)g.W * n >= ((size_t)1 << 31)) precomp = false;
  if (n == 0 || g.W > PRECOMP_MAX_W) precomp = false;
  if (!precomp && !(((unsigned)flags >> 8) & 0xff) && g.c > 16) {      // W separate bucket sets: the widths above 16 were
    g.c = 16; g.W = (256 + g.c - 1) / g.c; g.B = 1u << (g.c - 1);          // measured with ONE shared set (precomputed table)
  }
  g.precomp = precomp ? 1 : 0;
  g.nsets = precomp ? 1 : g.W;
  g.nb = g.nsets * g.B;
  size_t bytes = (precomp ? (size_t)g.W : 1) * (n ? n : 1) * 64;
  cudaError_t e = cudaMalloc(&b->pub.d_xy, bytes);
  if (e != cudaSuccess) { cudaGetLastError(); delete b; TRP_FAIL(ctx, TRP_E_OOM, "allocating %zu bytes for bases failed: %s", bytes, cudaGetErrorString(e)); }
  if (n) {
    e = cudaMemcpyAsync(b->pub.d_xy, src, n * 64, src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { cudaFree(b->pub.d_xy); delete b; TRP_FAIL(ctx, TRP_E_CUDA, "bases upload failed: %s", cudaGetErrorString(e)); }
    if (precomp) {
      unsigned blocks = (unsigned)((n + 127) / 128);
      if (base_field_of(ctx->curve) == 0) msm_precompute_kernel<FpParams><<<blocks, 128, 0, ctx->stream>>>((uint4*)b->pub.d_xy, n, g.c, g.W);
      else msm_precompute_kernel<FqParams><<<blocks, 128, 0, ctx->stream>>>((uint4*)b->pub.d_xy, n, g.c, g.W);
      ctx->launches++;
      e = cudaGetLastError();
      if (e != cudaSuccess) { cudaFree(b->pub.d_xy); delete b; TRP_FAIL(ctx, TRP_E_CUDA, "precompute launch failed: %s", cudaGetErrorString(e)); }
    }
    e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { cudaFree(b->pub.d_xy); delete b; TRP_FAIL(ctx, TRP_E_CUDA, "bases load failed: %s", cudaGetErrorString(e)); }
  }
  *out = &b->pub;
  return TRP_OK;
}

void trp_bases_destroy(trp_bases* b) {
  if (!b) return;
  trp_bases_impl* impl = reinterpret_cast<trp_bases_impl*>(b);
  if (b->d_xy) cudaFree(b->d_xy);
  delete impl;
}

int trp_bases_info(const trp_bases* b, unsigned* c, unsigned* W, unsigned* precomp) {
  const trp_bases_impl* impl = reinterpret_cast<const trp_bases_impl*>(b);
  if (c) *c = impl->g.c;
  if (W) *W = impl->g.W;
  if (precomp) *precomp = impl->g.precomp;
  return TRP_OK;
}

// ---- inclusive prefix sums of affine points: out[j] = P_0 + ... + P_j (affine) -------------------------------------------
// The base set of the "summation by parts" commitment (capi.cu, trp_dev_points_prefix_sum): a column z whose value rarely
// CHANGES from row to row (halo2's grand-product columns are constant over the unused rows of a circuit) commits as
//   sum_i z_i G_i = sum_j (z_j - z_{j+1}) Q_j,   Q_j = G_0 + ... + G_j,  z_n = 0,
// an MSM over Q with a SPARSE scalar column.  Q is built once per parameter set: per-thread chunk totals (PS_CH points), a
// two-level serial scan of the totals (segments of PS_SEG), then every thread replays its chunk from its offset and
// normalises its PS_CH results with one inversion.
constexpr unsigned PS_CH = 16, PS_SEG = 256;
template <class BPR>
__global__ void __launch_bounds__(128) points_chunk_totals_kernel(const uint4* in, size_t n, uint4* tot) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t start = t * PS_CH;
  if (start >= n) return;
  unsigned cnt = (unsigned)(n - start < PS_CH ? n - start : PS_CH);
  XYZZ<BPR> acc = xyzz_identity<BPR>();
  for (unsigned k = 0; k < cnt; ++k) xyzz_add_mixed(acc, load_affine<BPR>(in, start + k));
  store_xyzz(tot + 8 * t, acc);
}
// v[g * seg .. (g + 1) * seg) -> exclusive prefix sums within the segment, seg_tot[g] = the segment's total
template <class BPR>
__global__ void __launch_bounds__(64) points_scan_segments_kernel(uint4* v, size_t count, size_t seg, uint4* seg_tot) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t start = g * seg;
  if (start >= count) return;
  size_t end = start + seg < count ? start + seg : count;
  XYZZ<BPR> acc = xyzz_identity<BPR>();
  for (size_t i = start; i < end; ++i) {
    XYZZ<BPR> x = load_xyzz<BPR>(v + 8 * i);
    store_xyzz(v + 8 * i, acc);
    xyzz_add(acc, x);
  }
  store_xyzz(seg_tot + 8 * g, acc);
}
template <class BPR>
__global__ void __launch_bounds__(128) points_prefix_finish_kernel(const uint4* in, size_t n, const uint4* tot_excl, const uint4* seg_excl,
                                                                    const uint4* seg2_excl, uint4* out) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t start = t * PS_CH;
  if (start >= n) return;
  unsigned cnt = (unsigned)(n - start < PS_CH ? n - start : PS_CH);
  XYZZ<BPR> cur = load_xyzz<BPR>(seg2_excl + 8 * (t / PS_SEG / PS_SEG));
  xyzz_add(cur, load_xyzz<BPR>(seg_excl + 8 * (t / PS_SEG)));
  xyzz_add(cur, load_xyzz<BPR>(tot_excl + 8 * t));
  XYZZ<BPR> chain[PS_CH];
  Fe<BPR> pre[PS_CH];
  Fe<BPR> prod = fe_one<BPR>();
  for (unsigned k = 0; k < cnt; ++k) {
    xyzz_add_mixed(cur, load_affine<BPR>(in, start + k));
    chain[k] = cur;
    pre[k] = prod;
    if (!xyzz_is_identity(cur)) prod = fe_mul(prod, fe_mul(cur.zz, cur.zzz));
  }
  Fe<BPR> inv = fe_inv(prod);
  for (int k = (int)cnt - 1; k >= 0; --k) {
    uint4* slot = out + 4 * (start + k);
    const XYZZ<BPR> e = chain[k];
    if (xyzz_is_identity(e)) { for (int q = 0; q < 4; ++q) slot[q] = make_uint4(0, 0, 0, 0); continue; }
    Fe<BPR> einv = fe_mul(inv, pre[k]);
    inv = fe_mul(inv, fe_mul(e.zz, e.zzz));
    fe_store(slot, fe_mul(e.x, fe_mul(einv, e.zzz)));
    fe_store(slot + 2, fe_mul(e.y, fe_mul(einv, e.zz)));
  }
}

// d_in / d_out: n affine points (may alias: a thread reads its chunk before it writes it, and only its own chunk).
// ws: (ceil(n / PS_CH) + ceil(.. / PS_SEG) + ceil(.. / PS_SEG^2) + 1) XYZZ slots
size_t trp_points_prefix_ws_bytes(size_t n) {
  size_t t0 = (n + PS_CH - 1) / PS_CH, t1 = (t0 + PS_SEG - 1) / PS_SEG, t2 = (t1 + PS_SEG - 1) / PS_SEG;
  return (t0 + t1 + t2 + 1) * 128 + 1024;
}
int trp_points_prefix_sum_impl(trp_ctx* ctx, const void* d_in, size_t n, void* d_out, void* ws) {
  if (n == 0) return TRP_OK;
  const size_t t0 = (n + PS_CH - 1) / PS_CH, t1 = (t0 + PS_SEG - 1) / PS_SEG, t2 = (t1 + PS_SEG - 1) / PS_SEG;
  if (t2 > PS_SEG) TRP_FAIL(ctx, TRP_E_INVALID, "prefix sums of more than %zu points are not supported", (size_t)PS_CH * PS_SEG * PS_SEG * PS_SEG);
  uint4* tot = (uint4*)ws; uint4* seg = tot + 8 * t0; uint4* seg2 = seg + 8 * t1; uint4* top = seg2 + 8 * t2;
  auto launch = [&](auto tag) -> int {
    typedef decltype(tag) BPR;
    points_chunk_totals_kernel<BPR><<<(unsigned)((t0 + 127) / 128), 128, 0, ctx->stream>>>((const uint4*)d_in, n, tot);
    TRP_LAUNCHED(ctx);
    points_scan_segments_kernel<BPR><<<(unsigned)((t1 + 63) / 64), 64, 0, ctx->stream>>>(tot, t0, PS_SEG, seg);
    TRP_LAUNCHED(ctx);
    points_scan_segments_kernel<BPR><<<(unsigned)((t2 + 63) / 64), 64, 0, ctx->stream>>>(seg, t1, PS_SEG, seg2);
    TRP_LAUNCHED(ctx);
    points_scan_segments_kernel<BPR><<<1, 64, 0, ctx->stream>>>(seg2, t2, PS_SEG, top);     // t2 <= PS_SEG: one segment
    TRP_LAUNCHED(ctx);
    points_prefix_finish_kernel<BPR><<<(unsigned)((t0 + 127) / 128), 128, 0, ctx->stream>>>((const uint4*)d_in, n, tot, seg, seg2, (uint4*)d_out);
    TRP_LAUNCHED(ctx);
    return TRP_OK;
  };
  return base_field_of(ctx->curve) == 0 ? launch(FpParams()) : launch(FqParams());
}

int trp_points_progression_impl(trp_ctx* ctx, const uint64_t* p0, const uint64_t* d, size_t n, void* d_out) {
  if (n == 0) return TRP_OK;
  unsigned threads = (unsigned)((n + GEN_CH - 1) / GEN_CH);
  unsigned blocks = (threads + 127) / 128;
  auto launch = [&](auto tag) {
    typedef decltype(tag) BPR;
    Affine<BPR> a, b;
    for (int i = 0; i < 4; ++i) {
      a.x.v[2 * i] = (uint32_t)p0[i]; a.x.v[2 * i + 1] = (uint32_t)(p0[i] >> 32);
      a.y.v[2 * i] = (uint32_t)p0[4 + i]; a.y.v[2 * i + 1] = (uint32_t)(p0[4 + i] >> 32);
      b.x.v[2 * i] = (uint32_t)d[i]; b.x.v[2 * i + 1] = (uint32_t)(d[i] >> 32);
      b.y.v[2 * i] = (uint32_t)d[4 + i]; b.y.v[2 * i + 1] = (uint32_t)(d[4 + i] >> 32);
    }
    points_progression_kernel<BPR><<<blocks, 128, 0, ctx->stream>>>(a, b, n, (uint4*)d_out);
  };
  if (base_field_of(ctx->curve) == 0) launch(FpParams()); else launch(FqParams());
  TRP_LAUNCHED(ctx);
  return TRP_OK;
}

size_t trp_msm_ws_bytes(const trp_bases* bases, size_t n, size_t m) {
  return msm_ws_bytes(reinterpret_cast<const trp_bases_impl*>(bases)->g, n, m);
}

int trp_msm_impl(trp_ctx* ctx, const trp_bases* bases, const void* d_scalars, size_t n, size_t m, void* d_out_jac,
                 void* ws, size_t ws_bytes) {
  const trp_bases_impl* bs = reinterpret_cast<const trp_bases_impl*>(bases);
  if (n > bases->n) TRP_FAIL(ctx, TRP_E_INVALID, "MSM length %zu exceeds the %zu loaded bases", n, bases->n);
  if ((size_t)n * bs->g.W >= ((size_t)1 << 31)) TRP_FAIL(ctx, TRP_E_INVALID, "MSM of %zu points x %u windows exceeds the 2^31 entry limit", n, bs->g.W);
  if (ws_bytes < msm_ws_bytes(bs->g, n, m)) TRP_FAIL(ctx, TRP_E_INVALID, "internal: MSM workspace too small");
  if (m == 0) return TRP_OK;
  const bool pallas = ctx->curve == TRP_CURVE_PALLAS;
  if (n == 0) {
    TRP_CUDA(ctx, cudaMemsetAsync(d_out_jac, 0, m * 96, ctx->stream));   // identity: x = y = z = 0
    return TRP_OK;
  }
  const size_t mc = msm_cols_per_chunk(bs->g, n, m);
  for (size_t k = 0; k < m; k += mc) {
    size_t cols = m - k < mc ? m - k : mc;
    const uint4* sc = (const uint4*)d_scalars + 2 * k * n;
    uint4* out = (uint4*)d_out_jac + 6 * k;
    int rc = pallas ? msm_chunk<FpParams, FqParams>(ctx, bs, sc, n, cols, out, (char*)ws, ws_bytes)
                    : msm_chunk<FqParams, FpParams>(ctx, bs, sc, n, cols, out, (char*)ws, ws_bytes);
    if (rc != TRP_OK) return rc;
  }
  unsigned blocks = (unsigned)((m + 31) / 32);
  if (pallas) msm_normalize_kernel<FpParams><<<blocks, 32, 0, ctx->stream>>>((uint4*)d_out_jac, m);
  else msm_normalize_kernel<FqParams><<<blocks, 32, 0, ctx->stream>>>((uint4*)d_out_jac, m);
  TRP_LAUNCHED(ctx);
  return TRP_OK;
}

// ---- MSM over caller-owned bases (no table, one bucket set per window): the IPA rounds' <p'_hi, G'_lo> / <p'_lo, G'_hi>, whose
// bases change every round -------------------------------------------------------------------------
static trp_bases_impl transient_bases(trp_ctx* ctx, const void* d_xy, size_t n) {
  trp_bases_impl b;
  b.pub.ctx = ctx; b.pub.n = n; b.pub.d_xy = const_cast<void*>(d_xy);
  MsmGeom& g = b.g;
  g.c = choose_c(n ? n : 1);
  if (g.c > 16) g.c = 16;                 // W separate bucket sets here: the widths above 16 were measured with ONE shared set
  g.W = (256 + g.c - 1) / g.c;
  g.B = 1u << (g.c - 1);
  g.stride = n;
  g.precomp = 0;
  g.nsets = g.W;
  g.nb = g.nsets * g.B;
  return b;
}

size_t trp_msm_var_ws_bytes(size_t n, size_t m) {
  trp_bases_impl b = transient_bases(nullptr, nullptr, n);
  return msm_ws_bytes(b.g, n, m);
}

int trp_msm_var_impl(trp_ctx* ctx, const void* d_bases, const void* d_scalars, size_t n, size_t m, void* d_out_jac, void* ws, size_t ws_bytes) {
  trp_bases_impl b = transient_bases(ctx, d_bases, n);
  return trp_msm_impl(ctx, &b.pub, d_scalars, n, m, d_out_jac, ws, ws_bytes);
}

// ---- sum of a few group elements (combining the per-GPU partial sums of a point-range-split MSM) -----------------------
namespace {
template <class BPR>
__global__ void points_sum_kernel(const uint4* jac, size_t g, uint4* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  XYZZ<BPR> acc = xyzz_identity<BPR>();
  for (size_t i = 0; i < g; ++i) {
    Fe<BPR> X = fe_load<BPR>(jac + 6 * i), Y = fe_load<BPR>(jac + 6 * i + 2), Z = fe_load<BPR>(jac + 6 * i + 4);
    if (fe_is_zero(Z)) continue;
    XYZZ<BPR> q;                                   // Jacobian (X, Y, Z) -> XYZZ (X, Y, Z^2, Z^3)
    q.x = X; q.y = Y; q.zz = fe_sqr(Z); q.zzz = fe_mul(q.zz, Z);
    xyzz_add(acc, q);
  }
  Fe<BPR> X, Y, Z;
  if (xyzz_is_identity(acc)) { X = fe_zero<BPR>(); Y = X; Z = X; }
  else {
    Fe<BPR> inv = fe_inv(fe_mul(acc.zz, acc.zzz));
    X = fe_mul(acc.x, fe_mul(inv, acc.zzz)); Y = fe_mul(acc.y, fe_mul(inv, acc.zz)); Z = fe_one<BPR>();
  }
  fe_store(out, X); fe_store(out + 2, Y); fe_store(out + 4, Z);
}
}  // namespace

int trp_points_sum_impl(trp_ctx* ctx, const void* d_jac, size_t g, void* d_out) {
  if (base_field_of(ctx->curve) == 0) points_sum_kernel<FpParams><<<1, 32, 0, ctx->stream>>>((const uint4*)d_jac, g, (uint4*)d_out);
  else points_sum_kernel<FqParams><<<1, 32, 0, ctx->stream>>>((const uint4*)d_jac, g, (uint4*)d_out);
  TRP_LAUNCHED(ctx);
  return TRP_OK;
}
