// Host-side lowering of the PUBLIC quotient program (include/tr_prover.h: 13 opcodes, { op, dst, a, b }) into the form the
// kernel of quotient.cu executes.  Plain C++ (no CUDA) so that tests/qlower_host_shim.cpp can check it on the CPU box.
//
// Why: the public program is what a tree walk of halo2's poly::Ast emits (poly/evaluator.rs, halo2_proofs 0.2.0; the reference's
// gates are its input, e.g. /root/reference/src/circuits/tables/exe.rs:535-1080): a leaf is LOADed into a virtual register and
// consumed once by the operation above it.  Executed literally, a third of the instructions of the TinyRAM program are such
// loads (2744 of 8188), every one a dispatch plus a shared-memory round trip of the virtual register.  The lowering
//   1. folds a LOAD / CONST into its single consumer (the operand is then read straight from the column or the constant
//      table): `b` operand modes B_COL, B_CONST; the value of X on the coset (COSETX, one multiplication by zeta each time the
//      tree walk meets a LinearTerm: 188 times in the TinyRAM program) is computed once into an extra register;
//   2. turns the program into ACCUMULATOR form: the kernel keeps the result of the previous instruction in hardware registers
//      (ACC), operand a of every instruction IS the accumulator -- either left as it is (F_FWD_A: a was the previous result;
//      operands are swapped, SUB <-> RSUB, to bring a forwarded value to the a side) or first loaded from the register file --
//      and the result replaces it, so forwarding costs no register moves (on sm_100a a move is an IMAD.MOV on the same pipe as the
//      multiplier).  Results nobody reads from the register file afterwards are not stored (F_NOWB).
// The field operations performed per row are the same operations on the same values (additions and multiplications commute
// exactly, x + (-y) and x - y give the same canonical representative), so results are bit-identical.
//
// Lowered instruction = 4 x uint32:  x = kernel opcode (K_*) | flags << 5 | col << 11,  y = dst,  z = a,  w = b (register index,
// constant index, or the rotation of a column operand).  Constant index n_consts is zeta (appended by the library).
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace qlower {

enum Op : uint32_t { L_MOV = 0, L_ADD = 1, L_SUB = 2, L_MUL = 3, L_NEG = 4, L_DBL = 5, L_STORE = 6, L_RSUB = 7, L_NOP = 8 };   // RSUB: dst = b - a
enum BMode : uint32_t { B_REG = 0, B_CONST = 1, B_COL = 2, B_X = 3, B_A = 4, B_NONE = 5 };
enum Flag : uint32_t { F_FWD_A = 1, F_NOWB = 2, F_NO_A = 4 };
// what the kernel switches on: operation x source of operand b
enum KOp : uint32_t { K_NOP = 0, K_MOV_CONST, K_MOV_COL, K_MOV_X, K_MOV_REG, K_ADD_REG, K_ADD_CONST, K_ADD_COL, K_SUB_REG, K_SUB_CONST, K_SUB_COL,
                      K_RSUB_REG, K_RSUB_CONST, K_RSUB_COL, K_MUL_REG, K_MUL_CONST, K_MUL_COL, K_MUL_A, K_NEG, K_DBL, K_STORE, K_COUNT };
constexpr uint32_t MAX_COLS = 1u << 21;
constexpr int PAD = 2;                     // NOPs appended to a lowered program (the kernel fetches one instruction ahead)

struct Ins {
  uint32_t op, bm, fl, col, dst, a, b;
  bool dead;
};

inline uint32_t kop_of(const Ins& i) {
  switch (i.op) {
    case L_MOV: return i.bm == B_CONST ? K_MOV_CONST : i.bm == B_COL ? K_MOV_COL : i.bm == B_X ? K_MOV_X : K_MOV_REG;
    case L_ADD: return i.bm == B_CONST ? K_ADD_CONST : i.bm == B_COL ? K_ADD_COL : K_ADD_REG;
    case L_SUB: return i.bm == B_CONST ? K_SUB_CONST : i.bm == B_COL ? K_SUB_COL : K_SUB_REG;
    case L_RSUB: return i.bm == B_CONST ? K_RSUB_CONST : i.bm == B_COL ? K_RSUB_COL : K_RSUB_REG;
    case L_MUL: return i.bm == B_CONST ? K_MUL_CONST : i.bm == B_COL ? K_MUL_COL : i.bm == B_A ? K_MUL_A : K_MUL_REG;
    case L_NEG: return K_NEG;
    case L_DBL: return K_DBL;
    case L_STORE: return K_STORE;
    default: return K_NOP;
  }
}
inline uint32_t pack_x(const Ins& i) { return kop_of(i) | (i.fl << 5) | (i.col << 11); }

struct Stats { size_t in = 0, out = 0, fused = 0, fwd = 0, nowb = 0, hoisted_x = 0, negs = 0; };

inline bool reads_a(const Ins& i) { return !(i.fl & F_NO_A) && i.op != L_NOP; }
inline bool reads_reg_b(const Ins& i) { return i.bm == B_REG; }
inline bool writes(const Ins& i) { return i.op != L_STORE && i.op != L_NOP; }
inline bool touches(const Ins& i, uint32_t r) {
  return (reads_a(i) && i.a == r) || (reads_reg_b(i) && i.b == r) || (writes(i) && i.dst == r);
}
inline bool reads(const Ins& i, uint32_t r) { return (reads_a(i) && i.a == r) || (reads_reg_b(i) && i.b == r); }

// is register r dead after instruction j (the next instruction that touches it, if any, overwrites it without reading it)?
inline bool dead_after(const std::vector<Ins>& v, size_t j, uint32_t r) {
  for (size_t k = j + 1; k < v.size(); ++k) {
    if (v[k].dead) continue;
    if (reads(v[k], r)) return false;
    if (writes(v[k]) && v[k].dst == r) return true;
  }
  return true;
}

// `prog` must already have passed validate_program (opcodes and operand ranges).  Returns false if a column index does not fit.
// *n_regs_out = n_regs, or n_regs + 1 when the value of X (COSETX: one multiplication each) is used more than once: it is then
// computed once into an extra register that is never overwritten, and every COSETX becomes a read of it.
// COSETX (dst = zeta * (+-ext_omega^g)) becomes MOV dst <- X table; MUL dst <- dst * consts[n_consts] (= zeta).
inline bool lower(const uint32_t* prog, size_t n_instr_in, unsigned n_regs, size_t n_consts, std::vector<uint32_t>& out,
                  unsigned* n_regs_out, Stats* st = nullptr) {
  size_t n_x = 0;
  for (size_t i = 0; i < n_instr_in; ++i) n_x += prog[4 * i] == 8;
  const bool hoist = n_x >= 2;
  const uint32_t RX = n_regs, ZETA = (uint32_t)n_consts;
  *n_regs_out = n_regs + (hoist ? 1 : 0);
  std::vector<Ins> v;
  v.reserve(n_instr_in + n_x + 2);
  auto push_x = [&](uint32_t d) {
    v.push_back(Ins{L_MOV, B_X, F_NO_A, 0, d, 0, 0, false});
    v.push_back(Ins{L_MUL, B_CONST, 0, 0, d, d, ZETA, false});
  };
  if (hoist) push_x(RX);
  for (size_t i = 0; i < n_instr_in; ++i) {
    const uint32_t* q = prog + 4 * i;
    const uint32_t op = q[0], d = q[1], a = q[2], b = q[3];
    Ins x{L_NOP, B_NONE, F_NO_A, 0, d, a, b, false};
    if (op == 8 && !hoist) { push_x(d); continue; }
    switch (op) {
      case 0: if (a >= MAX_COLS) return false;
              x.op = L_MOV; x.bm = B_COL; x.col = a; x.a = 0; break;                    // LOAD: b = rotation
      case 1: x.op = L_MOV; x.bm = B_CONST; x.b = a; x.a = 0; break;                    // CONST
      case 2: x.op = L_ADD; x.bm = B_REG; x.fl = 0; break;
      case 3: x.op = L_SUB; x.bm = B_REG; x.fl = 0; break;
      case 4: x.op = L_MUL; x.bm = B_REG; x.fl = 0; break;
      case 5: x.op = L_NEG; x.fl = 0; x.b = 0; break;
      case 6: x.op = L_MUL; x.bm = B_A; x.fl = 0; x.b = 0; break;                       // SQR
      case 7: x.op = L_DBL; x.fl = 0; x.b = 0; break;
      case 8: x.a = 0; x.op = L_MOV; x.bm = B_REG; x.b = RX; break;                     // COSETX, hoisted: a copy of RX
      case 9: x.op = L_STORE; x.fl = F_NOWB; x.dst = 0; x.b = 0; break;                    // STORE a
      case 10: x.op = L_MUL; x.bm = B_CONST; x.fl = 0; break;
      case 11: x.op = L_ADD; x.bm = B_CONST; x.fl = 0; break;
      default: x.op = L_SUB; x.bm = B_CONST; x.fl = 0; break;
    }
    v.push_back(x);
  }
  const size_t n_instr = v.size();
  Stats s;
  s.in = n_instr_in;
  s.hoisted_x = hoist ? n_x : 0;
  // 0. x + (-y) -> x - y: halo2's `impl Sub for Ast` is self + (-other), so the tree walk emits NEG r; ADD d, x, r.  Both forms
  //    give the canonical representative of x - y, so the bits are the same.
  for (size_t i = 0; i < n_instr; ++i) {
    Ins& m = v[i];
    if (m.op != L_NEG || m.dst != m.a) continue;
    const uint32_t r = m.dst;
    size_t j = i + 1;
    while (j < n_instr && (v[j].dead || !touches(v[j], r))) ++j;
    if (j == n_instr) continue;
    Ins& c = v[j];
    if (c.op != L_ADD || c.bm != B_REG) continue;
    const bool as_a = c.a == r, as_b = c.b == r;
    if (as_a == as_b) continue;
    if (!(c.dst == r) && !dead_after(v, j, r)) continue;
    c.op = as_b ? L_SUB : L_RSUB;                                 // a + (-r) = a - r;  (-r) + b = b - r
    m.dead = true;
    ++s.negs;
  }
  // 1. fold a leaf (MOV from a column / constant / X) into its single consumer
  for (size_t i = 0; i < n_instr; ++i) {
    Ins& m = v[i];
    if (m.dead || m.op != L_MOV || (m.bm != B_COL && m.bm != B_CONST && m.bm != B_REG)) continue;
    const uint32_t r = m.dst;
    size_t j = i + 1;
    while (j < n_instr && (v[j].dead || !touches(v[j], r))) ++j;
    if (j == n_instr) continue;
    Ins& c = v[j];
    if (m.bm == B_REG) {                                          // a copy of RX (the only register moves there are): read RX instead
      if (!(writes(c) && c.dst == r) && !dead_after(v, j, r)) continue;
      bool any = false;
      if (reads_a(c) && c.a == r) { c.a = RX; any = true; }
      if (reads_reg_b(c) && c.b == r) { c.b = RX; any = true; }
      if (any) { m.dead = true; ++s.fused; }
      continue;
    }
    if (c.bm != B_REG || !reads_a(c) || (c.op != L_ADD && c.op != L_SUB && c.op != L_RSUB && c.op != L_MUL)) continue;
    const bool as_a = c.a == r, as_b = c.b == r;
    if (as_a == as_b) continue;                                   // read twice (or only overwritten): leave it
    if (!(writes(c) && c.dst == r) && !dead_after(v, j, r)) continue;   // the register is read again later
    if (as_a) {                                                   // bring the leaf to the b side
      if (c.op == L_SUB) c.op = L_RSUB;                           // r - b  ->  b' = leaf, a' = old b:  dst = leaf - a'
      else if (c.op == L_RSUB) c.op = L_SUB;
      c.a = c.b;
    }
    c.bm = m.bm; c.b = m.b; c.col = m.col;
    m.dead = true;
    ++s.fused;
  }
  // 2. accumulator form: an operand that is the previous result is brought to the a side and taken from the accumulator
  size_t prev = n_instr;
  for (size_t j = 0; j < n_instr; ++j) {
    if (v[j].dead) continue;
    if (prev != n_instr && writes(v[prev])) {
      const uint32_t d = v[prev].dst;
      Ins& c = v[j];
      const bool as_a = reads_a(c) && c.a == d, as_b = reads_reg_b(c) && c.b == d;
      if (as_b && !as_a && reads_a(c) && (c.op == L_ADD || c.op == L_MUL || c.op == L_SUB || c.op == L_RSUB)) {
        if (c.op == L_SUB) c.op = L_RSUB; else if (c.op == L_RSUB) c.op = L_SUB;
        const uint32_t t = c.a; c.a = c.b; c.b = t;                 // a (op) d  ->  d (op') a
      }
      if (reads_a(c) && c.a == d && !(reads_reg_b(c) && c.b == d)) {      // (both operands = d: read both from the register file)
        c.fl |= F_FWD_A;
        ++s.fwd;
        if ((writes(c) && c.dst == d) || dead_after(v, j, d)) { v[prev].fl |= F_NOWB; ++s.nowb; }
      }
    }
    prev = j;
  }
  out.clear();
  for (size_t i = 0; i < n_instr; ++i) {
    if (v[i].dead) continue;
    out.push_back(pack_x(v[i])); out.push_back(v[i].dst); out.push_back(v[i].a); out.push_back(v[i].b);
  }
  s.out = out.size() / 4;
  for (int pad = 0; pad < PAD; ++pad)                                                 // the kernel fetches up to PAD instructions ahead
    for (int k = 0; k < 4; ++k) out.push_back(k == 0 ? (uint32_t)K_NOP | ((F_NO_A | F_NOWB) << 5) : 0u);
  if (st) *st = s;
  return true;
}

// Device encoding of a lowered program for ONE launch: everything the kernel would otherwise compute per instruction is folded in
// here, on the host, once per call (the program is staged per call anyway) -- register indices become shared-memory slot offsets
// (reg << bd_log, 2^bd_log threads per CTA) and a column operand carries the column's device ADDRESS, so the kernel needs no
// pointer-table lookup (one dependent load and one level of prefetch less) and no index arithmetic for its register file:
//     x = kernel opcode | flags << 5 | (dst << bd_log) << 8      y = (a << bd_log) | rotation << 16 (int16)
//     z, w = b: register slot (z) / constant index (z) / column address (z = low, w = high 32 bits)
// Returns false if a rotation does not fit 16 bits or a register slot does not fit its field.
inline bool is_col(uint32_t k) { return k == K_MOV_COL || k == K_ADD_COL || k == K_SUB_COL || k == K_RSUB_COL || k == K_MUL_COL; }
inline bool is_regb(uint32_t k) { return k == K_MOV_REG || k == K_ADD_REG || k == K_SUB_REG || k == K_RSUB_REG || k == K_MUL_REG; }
inline bool stage(const std::vector<uint32_t>& low, unsigned n_regs, unsigned bd_log, const uint64_t* col_ptrs, size_t n_cols,
                  std::vector<uint32_t>& out) {
  if (((uint64_t)n_regs << bd_log) > 65536) return false;
  out.resize(low.size());
  for (size_t i = 0; i + 3 < low.size(); i += 4) {
    const uint32_t x = low[i], k = x & 31u, fl = (x >> 5) & 7u, col = x >> 11, dst = low[i + 1], a = low[i + 2], b = low[i + 3];
    uint32_t y = (fl & F_NO_A) ? 0u : (a << bd_log), z = 0, w = 0;
    if (is_col(k)) {
      const int32_t rot = (int32_t)b;
      if (rot < -32768 || rot > 32767 || col >= n_cols) return false;
      y |= ((uint32_t)rot & 0xffffu) << 16;
      z = (uint32_t)col_ptrs[col]; w = (uint32_t)(col_ptrs[col] >> 32);
    } else if (is_regb(k)) {
      z = b << bd_log;
    } else {
      z = b;                                                       // constant index (or unused)
    }
    out[i] = k | (fl << 5) | ((dst << bd_log) << 8); out[i + 1] = y; out[i + 2] = z; out[i + 3] = w;
  }
  return true;
}

}  // namespace qlower
