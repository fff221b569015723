"""CPU tests of csrc/qlower.h, the library-internal lowering of the public quotient program (include/tr_prover.h) into what
quotient_vm_kernel executes: leaf loads folded into their consumer, the value of X hoisted, results forwarded in hardware
registers, write-backs elided.  tests/qlower_host_shim.cpp runs the PUBLIC program and the LOWERED program with the product's
own field arithmetic (csrc/ff.cuh, host build); the two must agree bit for bit, and the public interpreter itself is held against
the oracle's (pasta_model.run_program, the restatement of poly::Evaluator::evaluate the GPU tests use)."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

import pasta_model as pm
import oracle as O
from ast_util import random_ast, gate_like_ast

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_qlower_host_shim.so")
R = 1 << 256


@pytest.fixture(scope="module")
def shim():
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", os.path.join(HERE, "qlower_host_shim.cpp"), "-o", SO])
    lib = ctypes.CDLL(SO)
    lib.qls_run.restype = ctypes.c_int
    return lib


@pytest.fixture(scope="module")
def P():
    import __graft_entry__ as ge
    ge.load_package()
    from tiny_ram_halo2_b200 import poly
    return poly


def _mont(vals, p):
    return O.ints_to_limbs([v * R % p for v in vals])


def _unmont(arr, p):
    Ri = pow(R, -1, p)
    return [v * Ri % p for v in O.limbs_to_ints(arr)]


def run_both(shim, prog, cols, xs, zeta, p):
    """cols: n_cols lists of `rows` canonical ints; xs: the value ext_omega^g per row (canonical); returns (public, lowered, stats)"""
    rows = len(xs)
    code = np.ascontiguousarray(prog.code, dtype=np.uint32)
    consts = _mont(prog.consts or [0], p)
    flat = _mont([v for c in cols for v in c] or [0], p)
    xraw, z = _mont(xs, p), _mont([zeta], p)
    out_pub, out_low = np.zeros((rows, 4), dtype=np.uint64), np.zeros((rows, 4), dtype=np.uint64)
    stats = np.zeros(8, dtype=np.uint64)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = shim.qls_run(vp(code), ctypes.c_size_t(len(code)), ctypes.c_uint(prog.n_regs), vp(consts), ctypes.c_size_t(len(prog.consts)), vp(flat), ctypes.c_size_t(len(cols)),
                      ctypes.c_size_t(rows), vp(xraw), vp(z), vp(out_pub), vp(out_low), vp(stats))
    assert rc == 0
    return out_pub, out_low, dict(zip(("in", "out", "fused", "fwd", "nowb", "hoisted_x", "regs", "negs"), (int(s) for s in stats)))


@pytest.mark.parametrize("seed", range(40))
def test_lowered_program_equals_public_program_on_random_asts(shim, P, seed):
    rng = random.Random(1000 + seed)
    F = pm.Fp
    k, j = rng.choice([(2, 3), (3, 4), (3, 6)])
    dom = pm.EvaluationDomain(F, j, k)
    rows = dom.extended_len()
    n_polys = rng.randrange(1, 6)
    polys = [[rng.randrange(F.p) for _ in range(rows)] for _ in range(n_polys)]
    ast = random_ast(P, rng, n_polys, depth=rng.randrange(1, 7), p=F.p)
    if seed % 3 == 0:                      # several LinearTerms: the hoisted X register
        ast = ast + P.LinearTerm(rng.randrange(F.p)) * (P.Poly(0) + P.LinearTerm(1)) - P.LinearTerm(5)
    prog = P.compile_ast(ast, F.p)
    xs = [pow(dom.extended_omega, g, F.p) for g in range(rows)]
    pub, low, st = run_both(shim, prog, polys, xs, dom.g_coset, F.p)
    assert (pub == low).all()
    assert st["out"] <= st["in"] + 1 and st["regs"] in (prog.n_regs, prog.n_regs + 1)
    # the shim's public interpreter against the oracle's (whole extended domain: rotation step = period, rows cyclic)
    period = 1 << (dom.extended_k - dom.k)
    stepped = P.Program(prog.code.copy(), prog.consts, prog.n_regs, prog.n_cols)
    loads = stepped.code[:, 0] == P.LOAD
    stepped.code[loads, 3] = (stepped.code[loads, 3].astype(np.int32) * period).astype(np.uint32)
    pub2, low2, _ = run_both(shim, stepped, polys, xs, dom.g_coset, F.p)
    assert (pub2 == low2).all()
    assert _unmont(pub2, F.p) == pm.run_program(dom, prog.code.tolist(), prog.consts, polys)


def test_gate_like_program_is_shortened(shim, P):
    F = pm.Fp
    rng = random.Random(3)
    rows = 16
    leaves = [P.Poly(i) for i in range(4)]
    ast = gate_like_ast(P, leaves, y=rng.randrange(F.p))
    for i in range(3):                      # three permutation-style factors (column + beta delta^i X + gamma)
        ast = ast * (leaves[i] + P.LinearTerm(rng.randrange(F.p)) + rng.randrange(F.p))
    prog = P.compile_ast(ast, F.p)
    polys = [[rng.randrange(F.p) for _ in range(rows)] for _ in range(4)]
    xs = [rng.randrange(F.p) for _ in range(rows)]
    pub, low, st = run_both(shim, prog, polys, xs, rng.randrange(F.p), F.p)
    assert (pub == low).all()
    c = prog.counts()
    assert st["hoisted_x"] == 3 and st["regs"] == prog.n_regs + 1
    assert st["fused"] >= c["load"] * 2 // 3 and st["out"] < st["in"] - st["fused"] + 2 and st["nowb"] > st["out"] // 3


def test_hand_written_corner_cases(shim, P):
    """programs a tree compiler never emits: a loaded register read twice, read after its consumer, overwritten unread, SUB with the
    leaf on the left, a result consumed as both operands, STORE in the middle, two STOREs"""
    F = pm.Fp
    L, C, ADD, SUB, MUL, NEG, SQR, DBL, X, ST, MULC, ADDC, SUBC = range(13)
    code = [
        (L, 0, 0, 0), (L, 1, 1, 1), (SUB, 2, 0, 1),               # r2 = c0 - c1[+1]; r0 and r1 stay live
        (MUL, 2, 2, 0),                                            # r0 read again
        (L, 3, 2, 0xffffffff), (SUB, 3, 3, 2),                     # leaf on the left of a SUB: r3 = c2[-1] - r2
        (L, 0, 1, 0),                                              # r0 overwritten
        (MUL, 1, 1, 1), (ADD, 1, 1, 1),                            # both operands the same register
        (ST, 0, 3, 0),
        (ADD, 3, 3, 0), (ADD, 3, 3, 1), (C, 2, 0, 0), (SUB, 3, 2, 3), (X, 2, 0, 0), (MUL, 3, 3, 2), (X, 0, 0, 0), (ADD, 3, 3, 0),
        (NEG, 3, 3, 0), (SQR, 3, 3, 0), (DBL, 3, 3, 0), (MULC, 1, 3, 1), (ADDC, 1, 1, 0), (SUBC, 1, 1, 1),
        (L, 2, 0, 2), (L, 0, 1, 0), (ADD, 0, 0, 0), (MUL, 1, 1, 2), (ADD, 1, 1, 0),
        (ST, 0, 1, 0),
    ]
    rng = random.Random(11)
    prog = P.Program(np.array(code, dtype=np.uint32), [rng.randrange(F.p), rng.randrange(F.p)], 4, 3)
    rows = 8
    polys = [[rng.randrange(F.p) for _ in range(rows)] for _ in range(3)]
    xs = [rng.randrange(F.p) for _ in range(rows)]
    pub, low, st = run_both(shim, prog, polys, xs, rng.randrange(F.p), F.p)
    assert (pub == low).all() and pub.any()


class _Captured(Exception):
    pass


def test_real_tinyram_program(shim, P):
    """the quotient program of the reference's TinyRamCircuit<8, 8> (the constraint system of W = 32 with smaller tables), captured
    from plonk.create_proof on the oracle's backend: identical results on random columns, and the shape of the saving the kernel's
    work model quotes (a third of the instructions are leaf loads; all but one of the X evaluations go)"""
    import sys
    import plonk_model as VM
    import tinyram_programs as TP
    from tiny_ram_halo2_b200 import plonk as PL, tinyram as TR
    T = sys.modules["tiny_ram_halo2_b200.trace"]
    C = pm.Vesta
    p = C.scalar.p
    circ, fixed, copies, adv, inst = TR.build(PL, TP.answer_only(T, 8), 6)
    got = {}

    class Capture(VM.PythonBackend):
        def quotient(self, ast, ext_polys):
            got["ast"], got["ncols"] = ast, len(ext_polys)
            raise _Captured()

    be = Capture(C, 6, circ.cs.degree())
    pk = PL.keygen(be, circ.cs, fixed, copies)
    rng = random.Random(5)
    with pytest.raises(_Captured):
        PL.create_proof(be, pk, inst, adv, lambda: rng.randrange(p), PL.Blake2bWrite(C.base.p, p))
    prog = got["ast"] if isinstance(got["ast"], P.Program) else P.compile_ast(got["ast"], p)
    assert prog.n_cols <= got["ncols"]
    rows = 8
    polys = [[rng.randrange(p) for _ in range(rows)] for _ in range(got["ncols"])]
    xs = [rng.randrange(p) for _ in range(rows)]
    pub, low, st = run_both(shim, prog, polys, xs, rng.randrange(p), p)
    assert (pub == low).all() and pub.any()
    c = prog.counts()
    n_x = int((prog.code[:, 0] == P.COSETX).sum())
    assert c["instr"] > 8000 and c["load"] > 2700 and n_x == len(circ.cs.permutation) == 188
    assert st["hoisted_x"] == n_x and st["regs"] == prog.n_regs + 1
    # 8188 -> 5883 instructions: 1732 of the 2744 leaf loads folded into their consumer (the rest feed a unary / constant operation
    # and are forwarded in hardware registers instead), 187 of the 188 multiplications by zeta gone, 387 negations folded into
    # subtractions; 4 in 5 results never reach the shared-memory register file
    assert (st["in"], st["out"], st["fused"], st["negs"]) == (8188, 5883, 1920, 387) and st["nowb"] > 0.75 * st["out"]


@pytest.mark.parametrize("seed", range(200))
def test_lowering_preserves_arbitrary_valid_programs(shim, P, seed):
    """the C ABI accepts ANY well-formed program, not only what a tree compiler emits: random straight-line programs over 2-6
    registers with every opcode, registers reused / overwritten unread / read many times, several STOREs (the last one wins per row),
    COSETX anywhere -- the lowered program must store the same values.  Only programs that read a register after writing it are
    generated (reading an unwritten virtual register is undefined in the public format too)."""
    rng = random.Random(7000 + seed)
    F = pm.Fp
    L, C, ADD, SUB, MUL, NEG, SQR, DBL, X, ST, MULC, ADDC, SUBC = range(13)
    n_regs, n_cols, n_consts = rng.randrange(2, 7), rng.randrange(1, 5), rng.randrange(1, 4)
    written, code = set(), []
    for _ in range(rng.randrange(5, 120)):
        r = rng.random()
        dst = rng.randrange(n_regs)
        if not written or r < 0.25:
            op = rng.choice([L, L, L, C, X])
            code.append((op, dst, rng.randrange(n_cols) if op == L else rng.randrange(n_consts) if op == C else 0,
                         (rng.choice([0, 0, 1, -1, 2, -3]) & 0xffffffff) if op == L else 0))
            written.add(dst)
        elif r < 0.65:
            a, b = rng.choice(sorted(written)), rng.choice(sorted(written))
            code.append((rng.choice([ADD, SUB, MUL]), dst, a, b)); written.add(dst)
        elif r < 0.8:
            code.append((rng.choice([NEG, SQR, DBL]), dst, rng.choice(sorted(written)), 0)); written.add(dst)
        elif r < 0.93:
            code.append((rng.choice([MULC, ADDC, SUBC]), dst, rng.choice(sorted(written)), rng.randrange(n_consts))); written.add(dst)
        else:
            code.append((ST, 0, rng.choice(sorted(written)), 0))
    code.append((ST, 0, rng.choice(sorted(written)), 0))
    prog = P.Program(np.array(code, dtype=np.uint32), [rng.randrange(F.p) for _ in range(n_consts)], n_regs, n_cols)
    rows = 8
    polys = [[rng.randrange(F.p) for _ in range(rows)] for _ in range(n_cols)]
    xs = [rng.randrange(F.p) for _ in range(rows)]
    pub, low, st = run_both(shim, prog, polys, xs, rng.randrange(F.p), F.p)
    assert (pub == low).all()
