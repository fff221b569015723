// Host build of csrc/qlower.h (the lowering of the public quotient program into the kernel's internal form) with two
// interpreters over the real field arithmetic of csrc/ff.cuh: one of the PUBLIC program (include/tr_prover.h semantics), one
// of the LOWERED program in its staged encoding (qlower::stage) that follows quotient_vm_kernel statement for statement
// (accumulator, one case per operation and operand source, write-back elision, zeta as the constant after the caller's).  Built by tests/test_qlower_cpu.py; not part of the product.
#include "../tiny-ram-halo2_b200/csrc/ff.cuh"
#include "../tiny-ram-halo2_b200/csrc/qlower.h"
#include <cstring>
#include <vector>
using namespace ff;
typedef Fe<FpParams> F;

static F ld(const uint32_t* p) { F r; memcpy(r.v, p, 32); return r; }

extern "C" int qls_run(const uint32_t* prog, size_t n_instr, unsigned n_regs, const uint32_t* consts, size_t n_consts, const uint32_t* cols, size_t n_cols,
                       size_t rows, const uint32_t* xraw, const uint32_t* zeta_limbs, uint32_t* out_pub, uint32_t* out_low, uint64_t* stats) {
  const F zeta = ld(zeta_limbs);
  auto col_at = [&](uint32_t c, size_t row, int rot) { return ld(cols + 8 * ((size_t)c * rows + (size_t)(((long long)row + rot) % (long long)rows + rows) % rows)); };
  // public program
  {
    std::vector<F> r(n_regs);
    for (size_t row = 0; row < rows; ++row)
      for (size_t pc = 0; pc < n_instr; ++pc) {
        const uint32_t op = prog[4 * pc], d = prog[4 * pc + 1], a = prog[4 * pc + 2], b = prog[4 * pc + 3];
        switch (op) {
          case 0: r[d] = col_at(a, row, (int)b); break;
          case 1: r[d] = ld(consts + 8 * a); break;
          case 2: r[d] = fe_add(r[a], r[b]); break;
          case 3: r[d] = fe_sub(r[a], r[b]); break;
          case 4: r[d] = fe_mul(r[a], r[b]); break;
          case 5: r[d] = fe_neg(r[a]); break;
          case 6: r[d] = fe_sqr(r[a]); break;
          case 7: r[d] = fe_dbl(r[a]); break;
          case 8: r[d] = fe_mul(ld(xraw + 8 * row), zeta); break;
          case 9: memcpy(out_pub + 8 * row, r[a].v, 32); break;
          case 10: r[d] = fe_mul(r[a], ld(consts + 8 * b)); break;
          case 11: r[d] = fe_add(r[a], ld(consts + 8 * b)); break;
          default: r[d] = fe_sub(r[a], ld(consts + 8 * b)); break;
        }
      }
  }
  std::vector<uint32_t> low;
  unsigned regs2 = 0;
  qlower::Stats st;
  const size_t n_consts_ = n_consts;
  if (!qlower::lower(prog, n_instr, n_regs, n_consts_, low, &regs2, &st)) return 1;
  stats[0] = st.in; stats[1] = st.out; stats[2] = st.fused; stats[3] = st.fwd; stats[4] = st.nowb; stats[5] = st.hoisted_x; stats[6] = regs2; stats[7] = st.negs;
  const size_t n_low = low.size() / 4 - qlower::PAD;
  if (n_low != st.out) return 2;
  // the staged encoding the kernel reads (qlower::stage) with one "thread" per CTA (bd_log = 0: slot = register index) and the
  // column INDEX standing in for the column's device address
  std::vector<uint64_t> fake_ptrs(n_cols ? n_cols : 1);
  for (size_t c = 0; c < n_cols; ++c) fake_ptrs[c] = 0x100000000ull * (c + 1) + c;     // exercises both halves of the 64-bit field
  std::vector<uint32_t> sg;
  if (!qlower::stage(low, regs2, 0, fake_ptrs.data(), n_cols, sg)) return 3;
  {
    using namespace qlower;
    std::vector<F> r(regs2);
    auto cst = [&](uint32_t i) { return i == n_consts_ ? zeta : ld(consts + 8 * i); };      // the library appends zeta
    for (size_t row = 0; row < rows; ++row) {
      // poison the register file between rows: a lowered program must not depend on what an earlier row left behind
      for (auto& x : r) for (int i = 0; i < 8; ++i) x.v[i] = 0xdeadbeefu;
      F acc = fe_zero<FpParams>();
      for (size_t pc = 0; pc < n_low; ++pc) {
        const uint32_t x = sg[4 * pc], y = sg[4 * pc + 1], z = sg[4 * pc + 2], w = sg[4 * pc + 3];
        const uint32_t fl = x >> 5, dst = x >> 8;
        if (!(fl & (F_FWD_A | F_NO_A))) acc = r[y & 0xffffu];
        auto col = [&]() {
          const uint64_t ptr = ((uint64_t)w << 32) | z;
          const uint32_t c = (uint32_t)(ptr & 0xffffffffu);
          if (ptr != 0x100000000ull * (c + 1) + c) { F bad; for (int i = 0; i < 8; ++i) bad.v[i] = 0xbadbadu; return bad; }
          return col_at(c, row, (int)(int16_t)(y >> 16));
        };
        switch (x & 31u) {
          case K_MOV_CONST: acc = cst(z); break;
          case K_MOV_COL: acc = col(); break;
          case K_MOV_REG: acc = r[z]; break;
          case K_MOV_X: acc = ld(xraw + 8 * row); break;
          case K_ADD_REG: acc = fe_add(acc, r[z]); break;
          case K_ADD_CONST: acc = fe_add(acc, cst(z)); break;
          case K_ADD_COL: acc = fe_add(acc, col()); break;
          case K_SUB_REG: acc = fe_sub(acc, r[z]); break;
          case K_SUB_CONST: acc = fe_sub(acc, cst(z)); break;
          case K_SUB_COL: acc = fe_sub(acc, col()); break;
          case K_RSUB_REG: acc = fe_sub(r[z], acc); break;
          case K_RSUB_CONST: acc = fe_sub(cst(z), acc); break;
          case K_RSUB_COL: acc = fe_sub(col(), acc); break;
          case K_MUL_REG: acc = fe_mul(acc, r[z]); break;
          case K_MUL_CONST: acc = fe_mul(acc, cst(z)); break;
          case K_MUL_COL: acc = fe_mul(acc, col()); break;
          case K_MUL_A: acc = fe_mul(acc, acc); break;
          case K_NEG: acc = fe_neg(acc); break;
          case K_DBL: acc = fe_dbl(acc); break;
          case K_STORE: memcpy(out_low + 8 * row, acc.v, 32); break;
          default: break;
        }
        if (!(fl & F_NOWB)) r[dst] = acc;
      }
    }
  }
  return 0;
}
