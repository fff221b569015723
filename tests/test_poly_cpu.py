"""CPU tests of the host-side Ast -> program compiler (tiny-ram-halo2_b200/poly.py): the compiled program, run by the
oracle's reference interpreter, must equal the oracle's direct Ast evaluation (restatement of poly::Evaluator::evaluate)."""
import random

import numpy as np
import pytest

import pasta_model as pm
from ast_util import random_ast, gate_like_ast


@pytest.fixture(scope="module")
def P():
    import __graft_entry__ as ge
    ge.load_package()
    from tiny_ram_halo2_b200 import poly
    return poly


@pytest.mark.parametrize("seed", range(12))
def test_compiler_matches_ast_semantics(P, seed):
    rng = random.Random(seed)
    F = pm.Fp
    k, j = rng.choice([(2, 3), (3, 4), (3, 6), (2, 9)])
    dom = pm.EvaluationDomain(F, j, k)
    rows = dom.extended_len()
    n_polys = 4
    polys = [[rng.randrange(F.p) for _ in range(rows)] for _ in range(n_polys)]
    ast = random_ast(P, rng, n_polys, depth=rng.randrange(2, 6), p=F.p)
    prog = P.compile_ast(ast, F.p)
    assert prog.code[-1, 0] == P.STORE and prog.n_regs >= 1
    want = pm.evaluate_ast(dom, ast, polys)
    got = pm.run_program(dom, prog.code.tolist(), prog.consts, polys)
    assert got == want


def test_register_pressure_is_sethi_ullman(P):
    # a left-deep sum of 200 products needs 2 registers, a balanced product tree of 16 leaves needs 5
    leaves = [P.Poly(i) for i in range(16)]
    s = leaves[0] * leaves[1]
    for i in range(200):
        s = s + leaves[i % 16] * leaves[(i + 1) % 16]
    assert P.compile_ast(s, pm.Fp.p).n_regs == 3
    layer = leaves
    while len(layer) > 1:
        layer = [layer[i] * layer[i + 1] for i in range(0, len(layer), 2)]
    assert P.compile_ast(layer[0], pm.Fp.p).n_regs == 5


def test_coset_mode_matches_extended_mode(P):
    """Evaluating coset by coset (columns on the size-n coset, rotation step 1) gives the rows j, j+2^(ek-k), ... of the
    whole-domain evaluation -- the identity the device-resident prover relies on."""
    F = pm.Fp
    dom = pm.EvaluationDomain(F, 6, 3)
    rng = random.Random(5)
    period = 1 << (dom.extended_k - dom.k)
    coeffs = [[rng.randrange(F.p) for _ in range(dom.n)] for _ in range(4)]
    ext = [dom.coeff_to_extended(c) for c in coeffs]
    ast = gate_like_ast(P, [P.Poly(i) for i in range(4)], y=rng.randrange(F.p)) + P.LinearTerm(7)
    prog = P.compile_ast(ast, F.p)
    whole = pm.run_program(dom, prog.code.tolist(), prog.consts, ext)
    assert whole == pm.evaluate_ast(dom, ast, ext)
    for j in range(period):
        on_coset = [e[j::period] for e in ext]
        got = pm.run_program(dom, prog.code.tolist(), prog.consts, on_coset, coset=j)
        assert got == whole[j::period]


def test_quotient_identity(P):
    """h = (a*b - c) / (X^n - 1) computed through evaluate -> divide_by_vanishing_poly -> extended_to_coeff satisfies
    h(x) * (x^n - 1) = a(x) b(x) - c(x) at a random point (the check SURVEY.md 7 'hard parts' prescribes)."""
    F = pm.Fp
    dom = pm.EvaluationDomain(F, 3, 4)
    rng = random.Random(9)
    a_l = [rng.randrange(F.p) for _ in range(dom.n)]
    b_l = [rng.randrange(F.p) for _ in range(dom.n)]
    c_l = [x * y % F.p for x, y in zip(a_l, b_l)]
    co = [dom.lagrange_to_coeff(v) for v in (a_l, b_l, c_l)]
    ext = [dom.coeff_to_extended(v) for v in co]
    A, B, C = (P.Poly(i) for i in range(3))
    prog = P.compile_ast(A * B - C, F.p)
    num = pm.run_program(dom, prog.code.tolist(), prog.consts, ext)
    h = dom.extended_to_coeff(dom.divide_by_vanishing_poly(num))
    x = rng.randrange(F.p)
    lhs = pm.eval_polynomial(F, h, x) * (pow(x, dom.n, F.p) - 1) % F.p
    rhs = (pm.eval_polynomial(F, co[0], x) * pm.eval_polynomial(F, co[1], x) - pm.eval_polynomial(F, co[2], x)) % F.p
    assert lhs == rhs


def test_unregistered_poly_is_rejected(P):
    class FakeCtx:
        curve = 1
    ev = P.Evaluator(FakeCtx())
    ev.register_poly(np.zeros((8, 4), dtype=np.uint64))
    with pytest.raises(ValueError):
        ev.compile(P.Poly(0) * P.Poly(3))
