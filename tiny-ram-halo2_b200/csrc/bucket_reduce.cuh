// Two-level weighted bucket sum for the experimental reduction of msm.cu (TRP_MSM_REDUCE=2).
//
// The value of a bucket set is  sum_b (b + 1) B_b.  Cut into chunks of S = 2^log_chunk buckets, chunk ch starting at bucket
// ch * S contributes  acc_ch + ch * S * tot_ch  with  acc_ch = sum_k (k + 1) B_{ch S + k}  and  tot_ch = sum_k B_{ch S + k},  so
//     value = sum_ch acc_ch  +  S * sum_ch ch * tot_ch .
// The default reduction multiplies every chunk's total by its first bucket index (a doubling chain per chunk, as much work as
// the chunk's own running sums); here sum_ch ch * tot_ch is again a weighted sum, computed by running sums inside a thread's
// run of chunks and a (sum, weighted sum) tree across threads: one doubling chain per TREE LEVEL.
//
// Host/device code like ec.cuh, so tests/test_ff_host.py runs the same functions on the CPU box.
#pragma once
#include "ec.cuh"

namespace ec {

template <class PR> struct WNode {
  XYZZ<PR> a;   // sum of the chunks' own weighted sums acc_ch
  XYZZ<PR> s;   // sum of the chunk totals tot_ch
  XYZZ<PR> w;   // sum of (ch - first chunk of the node) * tot_ch
};

// a thread's m consecutive chunks; acc(ch), tot(ch) load chunk ch of the run
template <class PR, class LoadAcc, class LoadTot> FF_HD WNode<PR> wnode_leaf(unsigned m, LoadAcc acc, LoadTot tot) {
  WNode<PR> n;
  n.a = xyzz_identity<PR>(); n.s = xyzz_identity<PR>(); n.w = xyzz_identity<PR>();
  for (unsigned ch = m; ch-- > 1;) {       // tot(ch) enters s at step ch and s is added to w at steps ch .. 1: ch times
    xyzz_add(n.s, tot(ch));
    xyzz_add(n.w, n.s);
    xyzz_add(n.a, acc(ch));
  }
  if (m) { xyzz_add(n.s, tot(0)); xyzz_add(n.a, acc(0)); }
  return n;
}

// l covers 2^log_len chunks and r the chunks right after them: l <- l ++ r
template <class PR> FF_HD void wnode_combine(WNode<PR>& l, const WNode<PR>& r, unsigned log_len) {
  XYZZ<PR> t = r.s;
  for (unsigned i = 0; i < log_len; ++i) xyzz_dbl(t);
  xyzz_add(l.w, r.w);
  xyzz_add(l.w, t);
  xyzz_add(l.s, r.s);
  xyzz_add(l.a, r.a);
}

template <class PR> FF_HD XYZZ<PR> wnode_root(const WNode<PR>& n, unsigned log_chunk) {
  XYZZ<PR> t = n.w;
  for (unsigned i = 0; i < log_chunk; ++i) xyzz_dbl(t);
  XYZZ<PR> r = n.a;
  xyzz_add(r, t);
  return r;
}

}  // namespace ec
