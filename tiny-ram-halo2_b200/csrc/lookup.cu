// K9: lookup-argument permutation -- halo2_proofs::plonk::lookup::prover::permute_expression_pair
// (halo2_proofs 0.2.0 @ a95945254dcc, Cargo.lock:619-621; run once per lookup by Argument::commit_permuted inside
// create_proof, /root/reference/src/test_utils.rs:41,96; the reference declares 31 lookups: even_bits.rs:158-170,
// out_table.rs:33-74, shift.rs:142-165, circuits/mod.rs:52-57).  SURVEY.md 8(f) row f1.
//
// The CPU routine sorts the usable rows of the compressed input expression (Ord on the canonical integer), counts the
// table expression's values in a BTreeMap, puts every first occurrence of an input value into the permuted table at the
// same row and hands the left-over table values, ascending, to the repeated-input rows taken from the END.  The result
// is fully determined by the two multisets, so the GPU version is:
//   1. Montgomery -> canonical, LSD radix sort of both columns (8-bit digits, 256-bit keys; byte positions on which all
//      keys of both columns agree are skipped -- lookup values are small, typically 2..4 of the 32 passes remain);
//   2. flag first occurrences / repeated rows (input side) and consumed / left-over values (table side, binary search);
//   3. two exclusive scans rank the repeated rows and the left-overs; a scatter pairs left-over t with repeated row
//      (R - 1 - t), exactly the pop-from-the-end order of the reference;
//   4. canonical -> Montgomery on the way out.
// An input value that does not occur in the table (the reference returns Error::ConstraintSystemFailure) is reported
// through *all_found = 0.
#include "common.cuh"
#include "scan.cuh"

using namespace ff;

namespace {

constexpr int RS_THREADS = 256, RS_ITEMS = 8, RS_TILE = RS_THREADS * RS_ITEMS, RS_WARPS = RS_THREADS / 32;

struct Key { uint32_t v[8]; };
__device__ __forceinline__ Key key_load(const uint4* p) {
  uint4 lo = p[0], hi = p[1];
  Key k;
  k.v[0] = lo.x; k.v[1] = lo.y; k.v[2] = lo.z; k.v[3] = lo.w; k.v[4] = hi.x; k.v[5] = hi.y; k.v[6] = hi.z; k.v[7] = hi.w;
  return k;
}
__device__ __forceinline__ void key_store(uint4* p, const Key& k) {
  p[0] = make_uint4(k.v[0], k.v[1], k.v[2], k.v[3]);
  p[1] = make_uint4(k.v[4], k.v[5], k.v[6], k.v[7]);
}
__device__ __forceinline__ int key_cmp(const Key& a, const Key& b) {   // -1, 0, 1 on the 256-bit integers
#pragma unroll
  for (int i = 7; i >= 0; --i) {
    if (a.v[i] < b.v[i]) return -1;
    if (a.v[i] > b.v[i]) return 1;
  }
  return 0;
}
__device__ __forceinline__ unsigned key_digit(const Key& k, unsigned byte) { return (k.v[byte >> 2] >> ((byte & 3) * 8)) & 0xffu; }

// keys[a][i] = canonical(src_a[i]); diff |= keys ^ keys[0][0]  (which byte positions vary at all)
template <class PR>
__global__ void rs_prepare_kernel(const uint4* src0, const uint4* src1, size_t n, uint4* keys /* 2 x n */, uint32_t* diff /* 8 */) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint4* src = blockIdx.y ? src1 : src0;
  uint32_t d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (i < n) {
    Fe<PR> ref = fe_from_mont(fe_load<PR>(src0));
    Fe<PR> c = fe_from_mont(fe_load<PR>(src + 2 * i));
    fe_store(keys + 2 * ((size_t)blockIdx.y * n + i), c);
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] = c.v[k] ^ ref.v[k];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    uint32_t x = d[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x |= __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0 && x) atomicOr(diff + k, x);
  }
}

// counts[(a * 256 + digit) * ntiles + tile]
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const uint4* keys, size_t n, unsigned byte, unsigned ntiles, uint32_t* counts) {
  __shared__ uint32_t hist[256];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const uint4* src = keys + 2 * (size_t)blockIdx.y * n;
  const size_t base = (size_t)blockIdx.x * RS_TILE;
  const unsigned lane = threadIdx.x & 31;
#pragma unroll 1
  for (int r = 0; r < RS_ITEMS; ++r) {
    size_t i = base + (size_t)r * RS_THREADS + threadIdx.x;
    unsigned d = i < n ? key_digit(key_load(src + 2 * i), byte) : 0xffffffffu;
    unsigned peers = __match_any_sync(0xffffffffu, d);
    if (d != 0xffffffffu && (unsigned)(__ffs(peers) - 1) == lane) atomicAdd(&hist[d], __popc(peers));
  }
  __syncthreads();
  counts[((size_t)blockIdx.y * 256 + threadIdx.x) * ntiles + blockIdx.x] = hist[threadIdx.x];
}

// stable scatter: warp w of a tile owns the contiguous run [w * 256, (w + 1) * 256) of it, processed 32 keys per round in
// lane order, so (warp, round, lane) is the key's position in the tile
__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const uint4* keys, uint4* out, size_t n, unsigned byte, unsigned ntiles,
                                                                const uint32_t* offsets) {
  __shared__ uint32_t whist[RS_WARPS][256];
  for (int w = 0; w < RS_WARPS; ++w) whist[w][threadIdx.x] = 0;
  __syncthreads();
  const uint4* src = keys + 2 * (size_t)blockIdx.y * n;
  uint4* dst = out + 2 * (size_t)blockIdx.y * n;
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const size_t wbase = (size_t)blockIdx.x * RS_TILE + (size_t)wid * (RS_TILE / RS_WARPS);
  unsigned digit[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    size_t i = wbase + (size_t)r * 32 + lane;
    unsigned d = i < n ? key_digit(key_load(src + 2 * i), byte) : 0xffffffffu;
    unsigned peers = __match_any_sync(0xffffffffu, d);
    unsigned leader = __ffs(peers) - 1;
    unsigned old = 0;
    if (d != 0xffffffffu && leader == lane) { old = whist[wid][d]; whist[wid][d] = old + __popc(peers); }
    old = __shfl_sync(0xffffffffu, old, leader);
    digit[r] = d;
    rank[r] = old + __popc(peers & ((1u << lane) - 1));
    __syncwarp();
  }
  __syncthreads();
  {   // thread t turns the per-warp counts of digit t into the warps' global start positions
    uint32_t run = offsets[((size_t)blockIdx.y * 256 + threadIdx.x) * ntiles + blockIdx.x] - (uint32_t)((size_t)blockIdx.y * n);
    for (int w = 0; w < RS_WARPS; ++w) { uint32_t c = whist[w][threadIdx.x]; whist[w][threadIdx.x] = run; run += c; }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    if (digit[r] == 0xffffffffu) continue;
    size_t i = wbase + (size_t)r * 32 + lane;
    key_store(dst + 2 * (size_t)(whist[wid][digit[r]] + rank[r]), key_load(src + 2 * i));
  }
}

// largest-lower-bound search: does the sorted array hold `x`?
__device__ __forceinline__ bool sorted_contains(const uint4* arr, size_t n, const Key& x) {
  size_t lo = 0, hi = n;
  while (lo < hi) {
    size_t mid = (lo + hi) >> 1;
    int c = key_cmp(key_load(arr + 2 * mid), x);
    if (c == 0) return true;
    if (c < 0) lo = mid + 1; else hi = mid;
  }
  return false;
}

// blockIdx.y = 0: input side, flag[i] = 1 for REPEATED rows; first occurrences must be present in the table.
// blockIdx.y = 1: table side, flag[n + j] = 1 for LEFT-OVER values (everything but one instance of each value the input uses).
__global__ void lk_flags_kernel(const uint4* sa, const uint4* st, size_t n, uint32_t* flags, uint32_t* err) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4* self = blockIdx.y ? st : sa;
  const uint4* other = blockIdx.y ? sa : st;
  Key x = key_load(self + 2 * i);
  bool first = i == 0 || key_cmp(key_load(self + 2 * (i - 1)), x) != 0;
  bool in_other = first && sorted_contains(other, n, x);
  if (blockIdx.y == 0) {
    flags[i] = first ? 0u : 1u;
    if (first && !in_other) atomicOr(err, 1u);
  } else {
    flags[n + i] = (first && in_other) ? 0u : 1u;
  }
}

__global__ void lk_rows_kernel(const uint32_t* flags, const uint32_t* ranks, size_t n, uint32_t* rlist) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flags[i]) rlist[ranks[i]] = (uint32_t)i;
}

// ranks: exclusive scan of flags over [input | table] (2n + 1 entries): R = ranks[n], left-over rank of j = ranks[n + j] - R
template <class PR>
__global__ void lk_emit_kernel(const uint4* sa, const uint4* st, const uint32_t* flags, const uint32_t* ranks, const uint32_t* rlist,
                               size_t n, uint4* out_a, uint4* out_s, uint32_t* err) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t R = ranks[n];
  if (blockIdx.y == 0) {
    Fe<PR> a = fe_to_mont(fe_load<PR>(sa + 2 * i));
    fe_store(out_a + 2 * i, a);
    if (!flags[i]) fe_store(out_s + 2 * i, a);
  } else if (flags[n + i]) {
    uint32_t t = ranks[n + i] - R;
    if (t >= R) { atomicOr(err, 2u); return; }
    fe_store(out_s + 2 * (size_t)rlist[R - 1 - t], fe_to_mont(fe_load<PR>(st + 2 * i)));
  }
}

struct PairWs {
  uint4 *keys[2];
  uint32_t *counts, *offsets, *block_sums, *total, *flags, *ranks, *rlist, *diff, *err;
  size_t bytes;
};
PairWs carve_pair(size_t n, char* base) {
  WsCursor cur{base, 0, 0};
  PairWs w;
  size_t ntiles = (n + RS_TILE - 1) / RS_TILE;
  w.keys[0] = cur.take<uint4>(4 * (n ? n : 1));
  w.keys[1] = cur.take<uint4>(4 * (n ? n : 1));
  w.counts = cur.take<uint32_t>(2 * 256 * ntiles + 1);
  w.offsets = cur.take<uint32_t>(2 * 256 * ntiles + 1);
  w.block_sums = cur.take<uint32_t>(scan::SCAN_BLOCK);
  w.total = cur.take<uint32_t>(4);
  w.flags = cur.take<uint32_t>(2 * n + 1);
  w.ranks = cur.take<uint32_t>(2 * n + 2);
  w.rlist = cur.take<uint32_t>(n + 1);
  w.diff = cur.take<uint32_t>(16);      // diff[8] | err
  w.err = w.diff + 8;
  w.bytes = cur.off + 4096;
  return w;
}

template <class PR>
int permute_pair_run(trp_ctx* ctx, const void* d_input, const void* d_table, size_t n, void* d_pa, void* d_ps, void* ws, int* all_found) {
  if (n >= ((size_t)1 << 31)) TRP_FAIL(ctx, TRP_E_INVALID, "permute_expression_pair of %zu rows exceeds 2^31", n);
  if (all_found) *all_found = 1;
  if (n == 0) return TRP_OK;
  PairWs w = carve_pair(n, (char*)ws);
  const unsigned ntiles = (unsigned)((n + RS_TILE - 1) / RS_TILE);
  uint32_t h_diff[9];
  {
    ProfScope ps(ctx, PROF_LOOKUP_SORT);
    TRP_CUDA(ctx, cudaMemsetAsync(w.diff, 0, 16 * sizeof(uint32_t), ctx->stream));
    rs_prepare_kernel<PR><<<dim3((unsigned)((n + 255) / 256), 2), 256, 0, ctx->stream>>>((const uint4*)d_input, (const uint4*)d_table, n, w.keys[0], w.diff);
    TRP_LAUNCHED(ctx);
  }
  // the set of byte positions that need a pass decides the launch sequence: one 32-byte read-back
  TRP_CUDA(ctx, cudaMemcpyAsync(h_diff, w.diff, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  TRP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int cur = 0;
  {
    ProfScope ps(ctx, PROF_LOOKUP_SORT);
    for (unsigned byte = 0; byte < 32; ++byte) {
      if (((h_diff[byte >> 2] >> ((byte & 3) * 8)) & 0xffu) == 0) continue;
      rs_hist_kernel<<<dim3(ntiles, 2), RS_THREADS, 0, ctx->stream>>>(w.keys[cur], n, byte, ntiles, w.counts);
      TRP_LAUNCHED(ctx);
      // one scan over both columns' counters; column 1's positions start at n, which rs_scatter subtracts again
      TRP_TRY(scan::run_scan(ctx, w.counts, w.offsets, nullptr, w.block_sums, w.total, (size_t)2 * 256 * ntiles, 0, 1));
      rs_scatter_kernel<<<dim3(ntiles, 2), RS_THREADS, 0, ctx->stream>>>(w.keys[cur], w.keys[cur ^ 1], n, byte, ntiles, w.offsets);
      TRP_LAUNCHED(ctx);
      cur ^= 1;
    }
  }
  const uint4* sa = w.keys[cur];
  const uint4* st = w.keys[cur] + 2 * n;
  {
    ProfScope ps(ctx, PROF_PRODUCTS);
    dim3 grid((unsigned)((n + 127) / 128), 2);
    lk_flags_kernel<<<grid, 128, 0, ctx->stream>>>(sa, st, n, w.flags, w.err);
    TRP_LAUNCHED(ctx);
    TRP_TRY(scan::run_scan(ctx, w.flags, w.ranks, nullptr, w.block_sums, w.total, 2 * n, 0, 1));
    lk_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(w.flags, w.ranks, n, w.rlist);
    TRP_LAUNCHED(ctx);
    lk_emit_kernel<PR><<<grid, 128, 0, ctx->stream>>>(sa, st, w.flags, w.ranks, w.rlist, n, (uint4*)d_pa, (uint4*)d_ps, w.err);
    TRP_LAUNCHED(ctx);
  }
  TRP_CUDA(ctx, cudaMemcpyAsync(h_diff + 8, w.err, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  TRP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (all_found) *all_found = h_diff[8] == 0;
  return TRP_OK;
}

}  // namespace

size_t trp_permute_pair_ws_bytes(size_t rows) { return carve_pair(rows, nullptr).bytes; }

int trp_permute_pair_impl(trp_ctx* ctx, int field, const void* d_input, const void* d_table, size_t rows, void* d_perm_input,
                          void* d_perm_table, void* ws, int* all_found) {
  return field == 0 ? permute_pair_run<FpParams>(ctx, d_input, d_table, rows, d_perm_input, d_perm_table, ws, all_found)
                    : permute_pair_run<FqParams>(ctx, d_input, d_table, rows, d_perm_input, d_perm_table, ws, all_found);
}
