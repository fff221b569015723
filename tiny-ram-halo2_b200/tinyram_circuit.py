"""A SATISFIABLE constraint system with the shape of the reference's TinyRamCircuit<W, 8> and a witness
for it, generated on the device: the workload of a real plonk.create_proof at BASELINE.json's k = 20.

Shape, from the cited reference code:
  advice 263 = 94 program-table columns (tables/prog.rs:143) + 95 execution-table pc / program-line columns + 74 others
  (tables/exe.rs:540-552, even_bits.rs:98-99 x 14, ...); instance 94 (prog.rs:141); fixed 23 (time, pc, 18 table columns,
  3 selectors); 139 gate polynomials of degree <= 6 (circuits/sprod.rs:65-90 is the maximum); 31 lookups = 28 one-column
  even-bits lookups (even_bits.rs:158-165) + a 15-wide one (out_table.rs:33-74) + a 2-wide one (shift.rs:142-165) + the
  dynamic 95-wide lookup of the execution table's program line in the program table (circuits/mod.rs:52-57 -> prog.rs:170-192);
  equality on 188 columns (prog.rs:151-152) => 47 grand products at degree 6; rotations cur / next.

What is NOT the reference: the individual polynomial identities.  The reference's gates encode the TinyRAM semantics and its
witness comes from running a program (trace.rs, exe.rs:792-1080) -- code that stays on the Rust side of the boundary.  Here the
gates are stand-ins of the same count, degree and fan-in that a device-generated witness satisfies by construction:
  * "derived" gates   s * (d - f_1 f_2 ... f_m): the advice column d is computed from other columns (m <= 5, some rotated);
  * "boolean" gates   s * b * (b - 1) * (extra factors): b is one of the 92 selector columns of the looked-up program line;
  * lookups           inputs are drawn from the tables; the execution table's program line is a gather of program-table rows.
The scalar distribution follows SURVEY.md 8(a)'s note: the 2 x 92 selector columns are {0,1}, opcodes < 32, words < 2^16, a
few columns are full width."""
from __future__ import annotations

import random

import numpy as np

N_LINE = 94          # opcode, immediate, 92 selectors (prog.rs:65-77)
N_EVEN, N_OUT, N_SHIFT, N_UNIFORM, N_DERIVED = 28, 15, 2, 10, 19
N_BOOL_GATES_EXTRA = 28


def build(PL, be, seed: int = 40, scale: float = 1.0, program_len: int = None, copy_rows: int = 1024):
    """PL: the plonk module; be: plonk.GpuBackend.  Returns (cs, fixed, copies, advice, instances): fixed / advice / instances
    are device vectors (be.vec passthrough).  scale < 1 shrinks the column / gate / lookup counts for small-domain tests."""
    import torch
    ctx, lib, n, p = be.ctx, be.lib, be.n, be.p
    rng = random.Random(seed)
    sc = lambda x: max(1, int(round(x * scale)))
    n_line = max(3, sc(N_LINE))
    n_even, n_out, n_shift, n_uni, n_der = sc(N_EVEN), max(2, sc(N_OUT)), 2, max(2, sc(N_UNIFORM)), max(2, sc(N_DERIVED))
    A, F, I = PL.ADVICE, PL.FIXED, PL.INSTANCE
    cs = PL.ConstraintSystem()

    # ---- columns -----------------------------------------------------------------------------------------------------------------
    inst = [cs.instance_column() for _ in range(n_line)]
    prog = [cs.advice_column() for _ in range(n_line)]            # program table (dynamic lookup table, equality enabled)
    exe_pc = cs.advice_column()
    exe_line = [cs.advice_column() for _ in range(n_line)]
    even = [cs.advice_column() for _ in range(n_even)]
    out = [cs.advice_column() for _ in range(n_out)]
    shift = [cs.advice_column() for _ in range(n_shift)]
    uni = [cs.advice_column() for _ in range(n_uni)]
    der = [cs.advice_column() for _ in range(n_der)]
    f_time, f_pc, s_trace, s_table, s_prog = (cs.fixed_column() for _ in range(5))
    t_even = cs.fixed_column()
    t_out = [cs.fixed_column() for _ in range(n_out)]
    t_shift = [cs.fixed_column() for _ in range(n_shift)]
    q = cs.query

    # ---- sizes ----------------------------------------------------------------------------------------------------------------------
    # blinding_factors depends on the queries, which are fixed below; every advice column is queried at <= 2 rotations => 5
    bf = 5
    usable = n - (bf + 1)
    T = usable - 1                                                # rows the selectors enable (the last usable row stays free:
    #                                                               gates with a `next` rotation must not read a blinding row)
    TS = max(2, min(1 << 16, usable // 2))                        # table size (2^(W/2) rows for W = 32)
    L = min(program_len or (1 << 12), TS)                         # program length

    # ---- device helpers ----------------------------------------------------------------------------------------------------------
    gen = torch.Generator(device="cuda"); gen.manual_seed(seed)
    r2 = be._dev(be._limbs([be.R]))                               # R as a Montgomery element is R^2: small * R2 -> Montgomery

    def small(vals):
        """int64 tensor (n,) of values < 2^62 -> Montgomery column"""
        col = torch.zeros((n, 4), dtype=torch.int64, device="cuda")
        col[:, 0] = vals
        torch.cuda.synchronize()
        ctx.check(lib.trp_dev_field_op(ctx.handle, 0, 2 | 16, col.data_ptr(), r2.data_ptr(), col.data_ptr(), n))
        torch.cuda.synchronize()
        return col

    def op(code, a, b):
        o = torch.empty_like(a)
        torch.cuda.synchronize()
        ctx.check(lib.trp_dev_field_op(ctx.handle, 0, code, a.data_ptr(), b.data_ptr(), o.data_ptr(), n))
        torch.cuda.synchronize()
        return o

    mul = lambda a, b: op(2, a, b)
    rows = torch.arange(n, device="cuda", dtype=torch.int64)
    on = lambda limit: (rows < limit).to(torch.int64)
    randint = lambda hi: torch.randint(0, hi, (n,), device="cuda", dtype=torch.int64, generator=gen)

    def uniform():
        col = torch.empty((n, 4), dtype=torch.int64, device="cuda")
        col.random_(0, 1 << 62, generator=gen)                    # below 2^254: a valid Montgomery representation
        return col

    # ---- fixed columns ------------------------------------------------------------------------------------------------------------------
    fixed = [None] * cs.num_fixed
    fixed[f_time] = small(rows * on(T))
    fixed[f_pc] = small(rows * on(L))
    fixed[s_trace] = small(on(T)); fixed[s_table] = small(on(T)); fixed[s_prog] = small(on(L))
    fixed[t_even] = small(rows * on(TS))
    for i, c in enumerate(t_out):
        fixed[c] = small((i + 1) * rows * on(TS))
    fixed[t_shift[0]] = small(rows * on(TS))
    fixed[t_shift[1]] = small((7 * rows + 3) * on(TS) * (rows > 0))          # row 0 stays (0, 0): the image of disabled inputs

    # ---- witness -----------------------------------------------------------------------------------------------------------------------
    adv = [None] * cs.num_advice
    in_prog = on(L)
    prog_int = []
    for i in range(n_line):
        hi = 32 if i == 0 else (1 << 16) if i == 1 else 2          # opcode, immediate, boolean selectors
        v = randint(hi) * in_prog
        prog_int.append(v)
        adv[prog[i]] = small(v)
    t_row = randint(L) * on(T)                                     # the program line each execution row runs
    adv[exe_pc] = small(t_row)
    for i in range(n_line):
        adv[exe_line[i]] = small(prog_int[i][t_row] * on(T))
    even_int = [randint(TS) * on(T) for _ in range(n_even)]
    for c, v in zip(even, even_int):
        adv[c] = small(v)
    u = randint(TS) * on(T)
    for i, c in enumerate(out):
        adv[c] = small((i + 1) * u)
    v = randint(TS) * on(T)
    adv[shift[0]] = small(v)
    adv[shift[1]] = small((7 * v + 3) * (v > 0))
    for c in uni:
        adv[c] = uniform()

    # ---- gates ------------------------------------------------------------------------------------------------------------------------------
    free_cols = uni + even + [exe_pc] + exe_line[:2]
    for g in range(n_der):
        m = rng.choice([1, 2, 2, 3, 3, 4, 5])
        term_e, term_v = None, None
        for _ in range(m):
            c = rng.choice(free_cols)
            r = 1 if rng.random() < 0.25 else 0
            fe = q(A, c, r)
            fv = torch.roll(adv[c], -r, 0) if r else adv[c]
            term_e = fe if term_e is None else term_e * fe
            term_v = fv if term_v is None else mul(term_v, fv)
        adv[der[g]] = term_v
        cs.create_gate([q(F, s_trace) * (q(A, der[g]) - term_e)])
    sel_cols = exe_line[2:] if n_line > 2 else exe_line
    n_bool = len(sel_cols) + sc(N_BOOL_GATES_EXTRA)
    for g in range(n_bool):
        b = q(A, sel_cols[g % len(sel_cols)])
        poly = q(F, s_trace) * b * (b - 1)
        if g >= len(sel_cols):                                    # raise the degree with arbitrary extra factors (up to 6)
            for _ in range(rng.choice([1, 2, 3])):
                poly = poly * q(A, rng.choice(free_cols), rng.choice([0, 0, 1]))
        cs.create_gate([poly])

    # ---- lookups ------------------------------------------------------------------------------------------------------------------------------
    bsel = sel_cols[0]
    for c in even:
        cs.lookup([(q(F, s_table) * q(A, bsel) * q(A, c), q(F, t_even))])                     # input degree 3
    cs.lookup([(q(F, s_table) * q(A, c), q(F, t)) for c, t in zip(out, t_out)])
    cs.lookup([(q(F, s_table) * q(A, c), q(F, t)) for c, t in zip(shift, t_shift)])
    cs.lookup([(q(F, s_trace) * q(A, exe_pc), q(F, f_pc))] + [(q(F, s_trace) * q(A, e), q(A, t)) for e, t in zip(exe_line, prog)])

    # ---- equality: instance column i holds the program's column i -----------------------------------------------------------------------------
    for c in inst:
        cs.enable_equality(I, c)
    for c in prog:
        cs.enable_equality(A, c)
    copies = [((I, inst[i], r), (A, prog[i], r)) for i in range(n_line) for r in range(min(copy_rows, L))]
    instances = [small(v) for v in prog_int]
    assert cs.blinding_factors() == bf, "column queried at more rotations than assumed"
    return cs, fixed, copies, adv, instances
