"""GPU parity tests of the quotient evaluator (K6): trp_quotient_eval / trp_dev_quotient_eval / trp_dev_coeff_to_coset
through the C ABI against the oracle's restatement of poly::Evaluator::evaluate, bit-exact."""
import ctypes
import random

import numpy as np
import pytest

from util import O, pm
from ast_util import random_ast, gate_like_ast

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as ge
    return ge.load_package()


@pytest.fixture(scope="module")
def P(pkg):
    from tiny_ram_halo2_b200 import poly
    return poly


@pytest.fixture(scope="module")
def ctxs(pkg):
    return {O.VESTA: pkg.Context(0, pkg.VESTA), O.PALLAS: pkg.Context(0, pkg.PALLAS)}


def mont(field, ints):
    return O.to_mont(field, O.ints_to_limbs(ints))


def unmont(field, arr):
    return O.limbs_to_ints(O.from_mont(field, arr))


@pytest.mark.parametrize("curve", [O.VESTA, O.PALLAS])
@pytest.mark.parametrize("seed", range(10))
def test_evaluate_random_ast(pkg, P, ctxs, curve, seed):
    ctx = ctxs[curve]
    field = O.SCALAR_FIELD[curve]
    F = {O.FP: pm.Fp, O.FQ: pm.Fq}[field]
    rng = random.Random(100 + seed)
    k, j = rng.choice([(1, 3), (3, 4), (4, 6), (5, 6), (7, 3), (6, 9)])
    dom_m = pm.EvaluationDomain(F, j, k)
    dom = pkg.EvaluationDomain(ctx, j, k)
    rows = dom_m.extended_len()
    n_polys = 5
    polys = [[rng.randrange(F.p) for _ in range(rows)] for _ in range(n_polys)]
    ast = random_ast(P, rng, n_polys, depth=rng.randrange(2, 7), p=F.p)
    ev = P.new_evaluator(ctx)
    for v in polys:
        ev.register_poly(mont(field, v))
    got = unmont(field, ev.evaluate(ast, dom))
    assert got == pm.evaluate_ast(dom_m, ast, polys)


def test_many_registers_and_long_program(pkg, P, ctxs):
    """a balanced product tree (register-hungry) plus a 300-term y-fold (long program), k = 8"""
    ctx = ctxs[O.VESTA]
    F = pm.Fp
    rng = random.Random(3)
    dom_m = pm.EvaluationDomain(F, 3, 8)
    dom = pkg.EvaluationDomain(ctx, 3, 8)
    rows = dom_m.extended_len()
    polys = [[rng.randrange(F.p) for _ in range(rows)] for _ in range(8)]
    leaves = [P.Poly(i, rng.choice([0, 1, -1])) for i in range(8)] * 4
    layer = leaves
    while len(layer) > 1:
        layer = [layer[i] * layer[i + 1] + 3 for i in range(0, len(layer), 2)]
    y = rng.randrange(F.p)
    h = layer[0]
    for t in range(300):
        h = h * y + leaves[t % 32] * leaves[(t + 5) % 32]
    ev = P.new_evaluator(ctx)
    for v in polys:
        ev.register_poly(mont(O.FP, v))
    prog = ev.compile(h)
    assert prog.n_regs >= 6 and len(prog.code) > 1000
    assert unmont(O.FP, ev.evaluate(h, dom)) == pm.evaluate_ast(dom_m, h, polys)


@pytest.mark.parametrize("j,k", [(6, 4), (3, 6), (9, 5), (6, 11)])
def test_coeff_to_coset_and_coset_mode(pkg, P, ctxs, j, k):
    """device-resident path: coefficient columns -> per-coset evaluations (one size-n NTT each) -> VM per coset ->
    interleaved h_ext, equal to the whole-extended-domain evaluation and to the oracle."""
    import torch
    ctx = ctxs[O.VESTA]
    field = O.FP
    dom = pkg.EvaluationDomain(ctx, j, k)
    n, EN = 1 << k, dom.extended_len()
    period = EN // n
    ncols = 4
    coeff = O.random_field_mont(field, ncols * n, 70 + k).reshape(ncols, n, 4)
    ext = O.coeff_to_extended(field, j, k, coeff)                       # oracle, (ncols, EN, 4)
    y = 0x1234567890abcdef1234567890abcdef1234567890abcdef
    ast = gate_like_ast(P, [P.Poly(i) for i in range(4)], y) + P.LinearTerm(5)
    ev = P.new_evaluator(ctx)
    for c in range(ncols):
        ev.register_poly(ext[c])
    whole = ev.evaluate(ast, dom)
    if k <= 6:
        F = pm.Fp
        want = pm.evaluate_ast(pm.EvaluationDomain(F, j, k), ast, [unmont(field, ext[c]) for c in range(ncols)])
        assert unmont(field, whole) == want
    prog = ev.compile(ast)
    d_coeff = torch.from_numpy(coeff.view(np.int64)).cuda()
    d_coset = torch.empty_like(d_coeff)
    d_out = torch.zeros((EN, 4), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    for cs in range(period):
        ctx.check(ctx.lib.trp_dev_coeff_to_coset(dom.handle, d_coeff.data_ptr(), d_coset.data_ptr(), ncols, cs))
        ctx.sync()
        got_cols = d_coset.cpu().numpy().view(np.uint64)
        assert np.array_equal(got_cols, ext[:, cs::period, :]), cs
        ev.evaluate_device(prog, dom, [d_coset[c].data_ptr() for c in range(ncols)], d_out.data_ptr(), coset=cs)
        ctx.sync()
    assert np.array_equal(d_out.cpu().numpy().view(np.uint64), whole)
    # in-place coset transform (src == dst) gives the same columns
    d_inpl = d_coeff.clone()
    torch.cuda.synchronize()
    ctx.check(ctx.lib.trp_dev_coeff_to_coset(dom.handle, d_inpl.data_ptr(), d_inpl.data_ptr(), ncols, period - 1))
    ctx.sync()
    assert np.array_equal(d_inpl.cpu().numpy().view(np.uint64), ext[:, period - 1::period, :])


def test_quotient_identity_k12(pkg, P, ctxs):
    """h(x) (x^n - 1) = a(x) b(x) - c(x) at a random x for h computed entirely by the library (k = 12, j = 3)."""
    ctx = ctxs[O.VESTA]
    F, field = pm.Fp, O.FP
    k, j = 12, 3
    dom = pkg.EvaluationDomain(ctx, j, k)
    n = 1 << k
    a = O.random_field_mont(field, n, 1); b = O.random_field_mont(field, n, 2)
    c = ctx.field_op("mul", a, b)
    co = dom.lagrange_to_coeff(np.stack([a, b, c]))
    ext = dom.coeff_to_extended(co)
    ev = P.new_evaluator(ctx)
    A, B, C = (ev.register_poly(ext[i]) for i in range(3))
    num = ev.evaluate(A * B - C, dom)
    h = dom.extended_to_coeff(num, divide_by_vanishing_poly=True)
    x = 0x2b3c4d5e6f708192a3b4c5d6e7f8091a2b3c4d5e6f708192a3b4c5d6e7f8091 % F.p
    ci = [unmont(field, v) for v in co]
    hi = unmont(field, h)
    lhs = pm.eval_polynomial(F, hi, x) * (pow(x, n, F.p) - 1) % F.p
    rhs = (pm.eval_polynomial(F, ci[0], x) * pm.eval_polynomial(F, ci[1], x) - pm.eval_polynomial(F, ci[2], x)) % F.p
    assert lhs == rhs


def test_malformed_programs_are_rejected(pkg, P, ctxs):
    ctx = ctxs[O.VESTA]
    dom = pkg.EvaluationDomain(ctx, 3, 3)
    rows = dom.extended_len()
    col = np.zeros((rows, 4), dtype=np.uint64)
    out = np.zeros((rows, 4), dtype=np.uint64)
    colp = (ctypes.c_void_p * 1)(col.ctypes.data)
    consts = np.zeros((1, 4), dtype=np.uint64)

    def run(code, n_regs=2):
        code = np.array(code, dtype=np.uint32).reshape(-1, 4)
        return ctx.lib.trp_quotient_eval(dom.handle, code.ctypes.data_as(ctypes.c_void_p), len(code), n_regs,
                                         consts.ctypes.data_as(ctypes.c_void_p), 1, colp, 1, out.ctypes.data_as(ctypes.c_void_p))
    assert run([[P.LOAD, 0, 0, 0], [P.STORE, 0, 0, 0]]) == 0
    assert run([[P.LOAD, 0, 1, 0], [P.STORE, 0, 0, 0]]) == -1          # column out of range
    assert run([[P.LOAD, 5, 0, 0], [P.STORE, 0, 0, 0]]) == -1          # register out of range
    assert run([[P.CONST, 0, 3, 0], [P.STORE, 0, 0, 0]]) == -1         # constant out of range
    assert run([[99, 0, 0, 0], [P.STORE, 0, 0, 0]]) == -1              # unknown opcode
    assert run([[P.LOAD, 0, 0, 0]]) == -1                              # never stores
    assert run([[P.LOAD, 0, 0, 32767], [P.STORE, 0, 0, 0]]) == 0       # the widest rotation the staged encoding carries
    assert run([[P.LOAD, 0, 0, (-32768) & 0xffffffff], [P.STORE, 0, 0, 0]]) == 0
    assert run([[P.LOAD, 0, 0, 32768], [P.STORE, 0, 0, 0]]) == -1      # beyond it: rejected before anything is launched


@pytest.mark.parametrize("j,k", [(3, 4), (6, 5), (6, 11), (9, 6), (2, 3), (5, 12)])
def test_cosets_to_coeff(pkg, ctxs, j, k):
    """values of a degree < n(j-1) polynomial on the first j-1 cosets -> its coefficients (the reduced-coset form of
    extended_to_coeff . divide_by_vanishing_poly); the oracle's full-extended-domain evaluation supplies the coset values."""
    import torch
    ctx = ctxs[O.VESTA]
    field, F = O.FP, pm.Fp
    dom = pkg.EvaluationDomain(ctx, j, k)
    dom_m = pm.EvaluationDomain(F, j, k)
    n, EN, Q = 1 << k, dom.extended_len(), j - 1
    period = EN // n
    a = O.random_field_mont(field, Q * n, 80 + k)                         # coefficients of h, degree < Q n
    zeta_pows = np.stack([O.to_mont(field, O.ints_to_limbs([1]))[0], dom.g_coset, dom.g_coset_inv])
    scaled = ctx.field_op("mul", a, zeta_pows[np.arange(Q * n) % 3])
    padded = np.zeros((EN, 4), dtype=np.uint64); padded[:Q * n] = scaled
    h_ext = O.fft(field, padded, dom.extended_k, dom.extended_omega)      # h(zeta * ext_omega^r), oracle
    vals = np.stack([h_ext[i::period] for i in range(Q)])                 # (Q, n, 4)
    d_vals = torch.from_numpy(vals.view(np.int64)).cuda()
    d_out = torch.zeros((Q * n, 4), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctx.check(ctx.lib.trp_dev_cosets_to_coeff(dom.handle, d_vals.data_ptr(), Q, d_out.data_ptr(), 0))
    ctx.sync()
    assert np.array_equal(d_out.cpu().numpy().view(np.uint64), a)
    # numerator form: N = h * (X^n - 1), constant gamma_i - 1 on coset i; divide_by_vanishing undoes it
    num = vals.copy()
    for i in range(Q):
        gamma = pow(dom_m.g_coset * pow(dom_m.extended_omega, i, F.p) % F.p, n, F.p)
        c = np.tile(O.to_mont(field, O.ints_to_limbs([(gamma - 1) % F.p])), (n, 1))
        num[i] = ctx.field_op("mul", vals[i], c)
    d_vals = torch.from_numpy(num.view(np.int64)).cuda()
    torch.cuda.synchronize()
    ctx.check(ctx.lib.trp_dev_cosets_to_coeff(dom.handle, d_vals.data_ptr(), Q, d_out.data_ptr(), 1))
    ctx.sync()
    assert np.array_equal(d_out.cpu().numpy().view(np.uint64), a)
    # and it agrees with the full-domain route through the library: extended_to_coeff(h_ext)
    assert np.array_equal(dom.extended_to_coeff(h_ext), a)
    assert ctx.lib.trp_dev_cosets_to_coeff(dom.handle, d_vals.data_ptr(), period + 1, d_out.data_ptr(), 0) == -1


def test_reduced_coset_quotient_pipeline(pkg, P, ctxs):
    """device-resident quotient, reduced form: 5 of the 8 cosets (coeff_to_coset -> program with contiguous output) and
    cosets_to_coeff give the same h coefficients as evaluating the whole extended domain and calling extended_to_coeff."""
    import torch
    from tiny_ram_halo2_b200._lib import Q_CONTIGUOUS
    ctx = ctxs[O.VESTA]
    field, j, k = O.FP, 6, 7
    dom = pkg.EvaluationDomain(ctx, j, k)
    n, Q = 1 << k, j - 1
    lag = O.random_field_mont(field, 3 * n, 5).reshape(3, n, 4)
    lag[2] = ctx.field_op("mul", lag[0], lag[1])                          # c = a * b on every row => a*b - c vanishes on H
    sel = O.random_field_mont(field, n, 6)
    coeff = dom.lagrange_to_coeff(np.concatenate([lag, sel.reshape(1, n, 4)]))
    A, B, C, S = (P.Poly(i) for i in range(4))
    ast = S * S * S * (A * B - C) * (A.with_rotation(1) + 7)              # degree 6 numerator, divisible by X^n - 1
    ev = P.new_evaluator(ctx)
    ext = dom.coeff_to_extended(coeff)
    for c in range(4):
        ev.register_poly(ext[c])
    want = dom.extended_to_coeff(ev.evaluate(ast, dom), divide_by_vanishing_poly=True)
    prog = ev.compile(ast)
    d_coeff = torch.from_numpy(coeff.view(np.int64)).cuda()
    d_coset = torch.empty_like(d_coeff)
    d_vals = torch.zeros((Q, n, 4), dtype=torch.int64, device="cuda")
    d_out = torch.zeros((Q * n, 4), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    for cs in range(Q):
        ctx.check(ctx.lib.trp_dev_coeff_to_coset(dom.handle, d_coeff.data_ptr(), d_coset.data_ptr(), 4, cs))
        ev.evaluate_device(prog, dom, [d_coset[c].data_ptr() for c in range(4)], d_vals[cs].data_ptr(), coset=cs | Q_CONTIGUOUS)
    ctx.check(ctx.lib.trp_dev_cosets_to_coeff(dom.handle, d_vals.data_ptr(), Q, d_out.data_ptr(), 1))
    ctx.sync()
    assert np.array_equal(d_out.cpu().numpy().view(np.uint64), want)
