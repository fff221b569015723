"""GPU parity for SURVEY.md 8(f) row f1 -- batch inversion, grand products, the permutation / lookup Z columns and
permute_expression_pair -- through the C ABI, bit-exact against the oracle's restatement (oracle/pasta_model.py), plus
size-independent properties at BASELINE's k = 20."""
import random

import numpy as np
import pytest

from util import O, pm

pytestmark = pytest.mark.gpu

FIELD_OF = {O.VESTA: (O.FP, pm.Fp), O.PALLAS: (O.FQ, pm.Fq)}


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as ge
    return ge.load_package()


@pytest.fixture(scope="module")
def ctxs(pkg):
    return {O.VESTA: pkg.Context(0, pkg.VESTA), O.PALLAS: pkg.Context(0, pkg.PALLAS)}


def mont(field, ints):
    return O.to_mont(field, O.ints_to_limbs(ints))


def ints(field, arr):
    return O.limbs_to_ints(O.from_mont(field, arr))


@pytest.mark.parametrize("curve", [O.VESTA, O.PALLAS])
@pytest.mark.parametrize("n", [1, 7, 255, 256, 2047, 2048, 2049, 5000, 70001])
def test_batch_invert(ctxs, curve, n):
    ctx = ctxs[curve]
    field, F = FIELD_OF[curve]
    rng = random.Random(n)
    vals = [rng.randrange(F.p) for _ in range(n)]
    for i in range(0, n, 97):
        vals[i] = 0
    if n > 300:
        vals[256:290] = [0] * 34          # a whole thread's worth of zeros
    a = mont(field, vals)
    ctx.check(ctx.lib.trp_batch_invert(ctx.handle, 0, a.ctypes.data, n))
    assert ints(field, a) == pm.batch_invert(F, vals)


def test_batch_invert_all_zero_and_base_field(ctxs):
    ctx = ctxs[O.VESTA]
    a = np.zeros((3000, 4), dtype=np.uint64)
    ctx.check(ctx.lib.trp_batch_invert(ctx.handle, 0, a.ctypes.data, 3000))
    assert not a.any()
    vals = [5, 0, pm.Fq.p - 1, 123456789]
    b = mont(O.FQ, vals)                    # which_field = 1: the base field of Vesta is Fq
    ctx.check(ctx.lib.trp_batch_invert(ctx.handle, 1, b.ctypes.data, 4))
    assert ints(O.FQ, b) == pm.batch_invert(pm.Fq, vals)


@pytest.mark.parametrize("curve", [O.VESTA, O.PALLAS])
@pytest.mark.parametrize("n_in,n_out", [(0, 1), (1, 2), (9, 9), (2048, 2049), (2049, 2049), (6000, 6001), (50000, 49990)])
def test_grand_product(ctxs, curve, n_in, n_out):
    ctx = ctxs[curve]
    field, F = FIELD_OF[curve]
    rng = random.Random(n_in * 3 + n_out)
    vals = [rng.randrange(F.p) for _ in range(n_in)]
    init = rng.randrange(F.p)
    v = mont(field, vals) if n_in else np.zeros((1, 4), dtype=np.uint64)
    z = np.empty((n_out, 4), dtype=np.uint64)
    ctx.check(ctx.lib.trp_grand_product(ctx.handle, 0, v.ctypes.data, n_in, mont(field, [init]).ctypes.data, z.ctypes.data, n_out))
    want, run = [], init
    for i in range(n_out):
        want.append(run)
        if i < n_in:
            run = run * vals[i] % F.p
    assert ints(field, z) == want
    # init = NULL means 1
    ctx.check(ctx.lib.trp_grand_product(ctx.handle, 0, v.ctypes.data, n_in, None, z.ctypes.data, n_out))
    inv_init = pow(init, -1, F.p)
    assert ints(field, z) == [w * inv_init % F.p for w in want]


def test_grand_product_rejects_too_many_outputs(pkg, ctxs):
    ctx = ctxs[O.VESTA]
    v = np.zeros((4, 4), dtype=np.uint64); z = np.zeros((8, 4), dtype=np.uint64)
    with pytest.raises(pkg.TrpError):
        ctx.check(ctx.lib.trp_grand_product(ctx.handle, 0, v.ctypes.data, 4, None, z.ctypes.data, 6))


@pytest.mark.parametrize("curve", [O.VESTA, O.PALLAS])
@pytest.mark.parametrize("k,m,chunk,bf", [(1, 1, 1, 0), (4, 3, 4, 5), (6, 9, 4, 5), (11, 5, 2, 3), (12, 17, 16, 5)])
def test_permutation_commit(pkg, ctxs, curve, k, m, chunk, bf):
    ctx = ctxs[curve]
    field, F = FIELD_OF[curve]
    rng = random.Random(k * 100 + m)
    n = 1 << k
    vals = [[rng.randrange(F.p) for _ in range(n)] for _ in range(m)]
    sig = [[rng.randrange(F.p) for _ in range(n)] for _ in range(m)]
    if n > 8:
        vals[0][3] = 0; sig[0][5] = 0
    beta, gamma = rng.randrange(F.p), rng.randrange(F.p)
    draws = [rng.randrange(F.p) for _ in range(bf * ((m + chunk - 1) // chunk))]
    it1, it2 = iter(draws), iter(draws)
    dom = pkg.EvaluationDomain(ctx, 6, k)
    got = pkg.permutation.commit(dom, np.stack([mont(field, c) for c in vals]), np.stack([mont(field, c) for c in sig]), beta, gamma,
                                 chunk, bf, lambda: next(it1))
    want = pm.permutation_commit(F, F.root_of_unity(k), n, vals, sig, beta, gamma, chunk, bf, lambda: next(it2))
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert ints(field, g) == w


def test_permutation_chunk_limits(pkg, ctxs):
    ctx = ctxs[O.VESTA]
    dom = pkg.EvaluationDomain(ctx, 6, 3)
    cols = np.zeros((17, 8, 4), dtype=np.uint64)
    with pytest.raises(pkg.TrpError):
        pkg.permutation.commit(dom, cols, cols, 1, 2, 17, 0, lambda: 0)


def _lookup_case(F, rng, usable, distinct, big=False):
    table = [rng.randrange(distinct) for _ in range(usable)]
    if big:
        table[:6] = [F.p - 1, F.p - 2, 1 << 200, (1 << 64) + 5, (1 << 128) - 1, rng.randrange(F.p)]
    inp = [rng.choice(table) for _ in range(usable)]
    return inp, table


@pytest.mark.parametrize("curve", [O.VESTA, O.PALLAS])
@pytest.mark.parametrize("usable,distinct,big", [(1, 1, False), (2, 2, False), (31, 4, False), (250, 16, True), (2048, 300, False),
                                                 (2049, 70000, True), (10000, 1 << 40, True), (5003, 2, False)])
def test_permute_expression_pair(pkg, ctxs, curve, usable, distinct, big):
    ctx = ctxs[curve]
    field, F = FIELD_OF[curve]
    rng = random.Random(usable + distinct)
    inp, table = _lookup_case(F, rng, usable, distinct, big)
    pa, ps = pkg.lookup.permute_expression_pair(ctx, mont(field, inp + [77, 78]), mont(field, table + [79, 80]), usable)
    wa, ws = pm.permute_expression_pair(F, inp, table, usable)
    assert ints(field, pa) == wa
    assert ints(field, ps) == ws


def test_permute_expression_pair_failure_and_empty(pkg, ctxs):
    ctx = ctxs[O.VESTA]
    with pytest.raises(pkg.lookup.ConstraintSystemFailure):
        pkg.lookup.permute_expression_pair(ctx, mont(O.FP, [5, 6, 6]), mont(O.FP, [5, 5, 7]), 3)
    pa, ps = pkg.lookup.permute_expression_pair(ctx, np.zeros((4, 4), np.uint64), np.zeros((4, 4), np.uint64), 0)
    assert pa.shape == (0, 4) and ps.shape == (0, 4)
    # identical columns of one value: every byte position is constant, no radix pass runs
    same = mont(O.FP, [9] * 100)
    pa, ps = pkg.lookup.permute_expression_pair(ctx, same, same, 100)
    assert np.array_equal(pa, same) and np.array_equal(ps, same)


@pytest.mark.parametrize("curve", [O.VESTA, O.PALLAS])
@pytest.mark.parametrize("k,bf", [(3, 2), (7, 5), (12, 5)])
def test_lookup_commit_product(pkg, ctxs, curve, k, bf):
    ctx = ctxs[curve]
    field, F = FIELD_OF[curve]
    rng = random.Random(k)
    n = 1 << k
    usable = n - (bf + 1)
    inp, table = _lookup_case(F, rng, usable, 50)
    wa, ws = pm.permute_expression_pair(F, inp, table, usable)
    pad = lambda v: v + [rng.randrange(F.p) for _ in range(n - usable)]
    inp, table, wa, ws = pad(inp), pad(table), pad(wa), pad(ws)
    beta, gamma = rng.randrange(F.p), rng.randrange(F.p)
    draws = [rng.randrange(F.p) for _ in range(bf)]
    it1, it2 = iter(draws), iter(draws)
    dom = pkg.EvaluationDomain(ctx, 6, k)
    got = pkg.lookup.commit_product(dom, mont(field, inp), mont(field, table), mont(field, wa), mont(field, ws), beta, gamma, bf,
                                    lambda: next(it1))
    want = pm.lookup_commit_product(F, n, inp, table, wa, ws, beta, gamma, bf, lambda: next(it2))
    assert ints(field, got) == want
    assert want[usable] == 1


def test_lookup_argument_closes_at_k20(pkg, ctxs):
    """BASELINE size: permute 2^20 - 6 rows of a TinyRAM-like lookup (16-bit table values, skewed inputs) on the GPU and
    check the properties the verifier enforces: same multisets, a'[i] in {s'[i], a'[i-1]}, grand product back to 1."""
    ctx = ctxs[O.VESTA]
    k, bf = 20, 5
    n = 1 << k
    usable = n - (bf + 1)
    rng = np.random.Generator(np.random.PCG64(7))
    tab = np.zeros((n, 4), dtype=np.uint64); tab[:, 0] = rng.integers(0, 1 << 16, n, dtype=np.uint64)
    tab[:70000, 0] = np.arange(70000) % (1 << 16)          # every 16-bit value occurs
    inp = np.zeros((n, 4), dtype=np.uint64); inp[:, 0] = rng.integers(0, 1 << 8, n, dtype=np.uint64) ** 2
    a, s = O.to_mont(O.FP, inp), O.to_mont(O.FP, tab)
    pa, ps = pkg.lookup.permute_expression_pair(ctx, a, s, usable)
    ca, cs = O.from_mont(O.FP, pa), O.from_mont(O.FP, ps)
    assert not ca[:, 1:].any() and not cs[:, 1:].any()
    assert np.array_equal(ca[:, 0], np.sort(inp[:usable, 0]))
    assert np.array_equal(np.sort(cs[:, 0]), np.sort(tab[:usable, 0]))
    ok = ca[:, 0] == cs[:, 0]
    ok[1:] |= ca[1:, 0] == ca[:-1, 0]
    assert ok.all() and ca[0, 0] == cs[0, 0]
    dom = pkg.EvaluationDomain(ctx, 6, k)
    tail = O.random_field_mont(O.FP, 2 * (n - usable), 9)
    pa_full = np.concatenate([pa, tail[:n - usable]]); ps_full = np.concatenate([ps, tail[n - usable:]])
    z = pkg.lookup.commit_product(dom, a, s, pa_full, ps_full, 0x1234567, 0x7654321, bf, lambda: 1)
    one = O.to_mont(O.FP, O.ints_to_limbs([1]))[0]
    assert np.array_equal(z[0], one) and np.array_equal(z[usable], one)
