"""GPU parity for SURVEY.md 8(f) row f2 -- eval_polynomial, compute_inner_product, kate_division, the IPA round kernels and a
whole poly::commitment::prover::create_proof -- through the C ABI, bit-exact against oracle/pasta_model.py."""
import random

import numpy as np
import pytest

from util import O, pm, make_points

pytestmark = pytest.mark.gpu

FIELD_OF = {O.VESTA: (O.FP, pm.Fp), O.PALLAS: (O.FQ, pm.Fq)}
CURVE_OF = {O.VESTA: pm.Vesta, O.PALLAS: pm.Pallas}


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


def pt_limbs(curve, P):
    """pasta_model affine point (canonical ints / None) -> (8,) Montgomery limbs of the C ABI"""
    bf = O.BASE_FIELD[curve]
    if P is None:
        return np.zeros(8, dtype=np.uint64)
    return mont(bf, [P[0], P[1]]).reshape(8)


def pt_of(curve, limbs):
    bf = O.BASE_FIELD[curve]
    if not np.asarray(limbs).any():
        return None
    x, y = ints(bf, np.asarray(limbs, dtype=np.uint64).reshape(2, 4))
    return (x, y)


@pytest.mark.parametrize("curve", [O.VESTA, O.PALLAS])
@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 255, 256, 257, 8191, 8192, 8193, 16385, 100001])
def test_eval_polynomial(pkg, ctxs, curve, n):
    ctx = ctxs[curve]
    field, F = FIELD_OF[curve]
    rng = random.Random(n)
    coeffs = [rng.randrange(F.p) for _ in range(n)]
    for x in (rng.randrange(F.p), 0, 1, F.p - 1):
        got = pkg.ipa.eval_polynomial(ctx, mont(field, coeffs), mont(field, [x])[0])
        assert ints(field, got.reshape(1, 4))[0] == pm.eval_polynomial(F, coeffs, x)


def test_eval_polynomials_batched_device(pkg, ctxs):
    import torch
    ctx = ctxs[O.VESTA]
    rng = random.Random(5)
    n, m, stride = 5000, 7, 5003
    cols = [[rng.randrange(pm.Fp.p) for _ in range(stride)] for _ in range(m)]
    d = torch.from_numpy(np.stack([mont(O.FP, c) for c in cols]).view(np.int64)).cuda()
    out = torch.zeros((m, 4), dtype=torch.int64, device="cuda")
    x = rng.randrange(pm.Fp.p)
    torch.cuda.synchronize()
    ctx.check(ctx.lib.trp_dev_eval_polynomials(ctx.handle, 0, d.data_ptr(), stride, n, m, mont(O.FP, [x])[0].ctypes.data, out.data_ptr()))
    ctx.sync()
    assert ints(O.FP, out.cpu().numpy().view(np.uint64)) == [pm.eval_polynomial(pm.Fp, c[:n], x) for c in cols]


@pytest.mark.parametrize("n", [1, 300, 8192, 8200, 40000])
def test_eval_polynomials_at_pointer_table(pkg, ctxs, n):
    """trp_dev_eval_polynomials_at: separately allocated polynomials (the openings of create_proof), one point"""
    import ctypes
    import torch
    ctx = ctxs[O.VESTA]
    rng = random.Random(n)
    m = 9
    cols = [[rng.randrange(pm.Fp.p) for _ in range(n)] for _ in range(m)]
    cols[3] = [0] * n
    bufs = [torch.from_numpy(mont(O.FP, c).view(np.int64)).cuda() for c in cols]
    pad = torch.zeros(77, dtype=torch.int64, device="cuda")          # keeps the allocations from being one strided block
    out = torch.zeros((m, 4), dtype=torch.int64, device="cuda")
    tab = (ctypes.c_void_p * m)(*[b.data_ptr() for b in bufs])
    for x in (rng.randrange(pm.Fp.p), 0, 1):
        torch.cuda.synchronize()
        ctx.check(ctx.lib.trp_dev_eval_polynomials_at(ctx.handle, 0, tab, n, m, mont(O.FP, [x])[0].ctypes.data, out.data_ptr()))
        ctx.sync()
        assert ints(O.FP, out.cpu().numpy().view(np.uint64)) == [pm.eval_polynomial(pm.Fp, c, x) for c in cols]
    assert ctx.lib.trp_dev_eval_polynomials_at(ctx.handle, 0, None, n, m, mont(O.FP, [1])[0].ctypes.data, out.data_ptr()) != 0


@pytest.mark.parametrize("curve", [O.VESTA, O.PALLAS])
@pytest.mark.parametrize("n", [0, 1, 255, 256, 4097, 300000])
def test_compute_inner_product(pkg, ctxs, curve, n):
    ctx = ctxs[curve]
    field, F = FIELD_OF[curve]
    a = O.random_field_mont(field, max(n, 1), n + 1)[:n]
    b = O.random_field_mont(field, max(n, 1), n + 2)[:n]
    got = pkg.ipa.compute_inner_product(ctx, a, b)
    assert ints(field, got.reshape(1, 4))[0] == pm.compute_inner_product(F, ints(field, a) if n else [], ints(field, b) if n else [])
    with pytest.raises(ValueError):
        pkg.ipa.compute_inner_product(ctx, np.zeros((3, 4), np.uint64), np.zeros((4, 4), np.uint64))


@pytest.mark.parametrize("curve", [O.VESTA, O.PALLAS])
@pytest.mark.parametrize("n", [2, 3, 32, 33, 2048, 2049, 4097, 70000])
def test_kate_division(pkg, ctxs, curve, n):
    ctx = ctxs[curve]
    field, F = FIELD_OF[curve]
    rng = random.Random(n)
    a = [rng.randrange(F.p) for _ in range(n)]
    for b in (rng.randrange(F.p), 0, 1):
        got = pkg.ipa.kate_division(ctx, mont(field, a), mont(field, [b])[0])
        assert ints(field, got) == pm.kate_division(F, a, b)


def test_kate_division_identity_at_k20(pkg, ctxs):
    """(X - b) q(X) + p(b) = p(X), checked at a random point for a 2^20-coefficient polynomial."""
    ctx = ctxs[O.VESTA]
    F, n = pm.Fp, 1 << 20
    a = O.random_field_mont(O.FP, n, 77)
    b, r = 0x1234567890abcdef1234567890abcdef, 0xfedcba0987654321
    q = pkg.ipa.kate_division(ctx, a, mont(O.FP, [b])[0])
    ev = lambda poly, x: ints(O.FP, pkg.ipa.eval_polynomial(ctx, poly, mont(O.FP, [x])[0]).reshape(1, 4))[0]
    assert ((r - b) * ev(q, r) + ev(a, b)) % F.p == ev(a, r)


@pytest.mark.parametrize("curve", [O.VESTA, O.PALLAS])
def test_round_kernels(pkg, ctxs, curve):
    """powers, fold, parallel_generator_collapse and the MSM over caller-owned bases, as one IPA round uses them."""
    import torch
    ctx = ctxs[curve]
    field, F = FIELD_OF[curve]
    C = CURVE_OF[curve]
    rng = random.Random(11)
    half = 37
    x, u = rng.randrange(F.p), rng.randrange(F.p)
    d_b = torch.zeros((2 * half, 4), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctx.check(ctx.lib.trp_dev_powers(ctx.handle, 0, mont(field, [x])[0].ctypes.data, 2 * half, d_b.data_ptr()))
    ctx.check(ctx.lib.trp_dev_fold(ctx.handle, 0, d_b.data_ptr(), half, mont(field, [u])[0].ctypes.data))
    ctx.sync()
    pw = [pow(x, i, F.p) for i in range(2 * half)]
    assert ints(field, d_b.cpu().numpy().view(np.uint64)[:half]) == [(pw[i] + pw[i + half] * u) % F.p for i in range(half)]
    # generator collapse, with an identity in each half and a pair that cancels (g_lo = -[u] g_hi)
    G = (C.base.p - 1, 2)
    g = [C.mul(rng.randrange(1, F.p), G) for _ in range(2 * half)]
    g[3] = None; g[half + 5] = None
    g[7] = C.neg(C.mul(u, g[half + 7]))
    d_g = torch.from_numpy(np.stack([pt_limbs(curve, P) for P in g]).view(np.int64)).cuda()
    torch.cuda.synchronize()
    ctx.check(ctx.lib.trp_dev_generator_collapse(ctx.handle, d_g.data_ptr(), half, mont(field, [u])[0].ctypes.data))
    ctx.sync()
    got = [pt_of(curve, r) for r in d_g.cpu().numpy().view(np.uint64)[:half]]
    want = pm.parallel_generator_collapse(C, g, u)
    assert got == want and want[7] is None
    # MSM over the collapsed (caller-owned) bases
    sc = [rng.randrange(F.p) for _ in range(half)]
    d_sc = torch.from_numpy(mont(field, sc).view(np.int64)).cuda()
    d_out = torch.zeros(12, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctx.check(ctx.lib.trp_dev_msm_var(ctx.handle, d_g.data_ptr(), d_sc.data_ptr(), half, 1, d_out.data_ptr()))
    ctx.sync()
    assert pt_of(curve, O.jacobian_to_affine(curve, d_out.cpu().numpy().view(np.uint64).reshape(3, 4))) == C.best_multiexp(sc, want)


@pytest.mark.parametrize("n", [1 << 12, (1 << 14) + 1])
def test_msm_var_matches_oracle(pkg, ctxs, n):
    import torch
    ctx = ctxs[O.VESTA]
    pts = make_points(O.VESTA, n)
    sc = O.random_field_mont(O.FP, n, 3)
    d_pts = torch.from_numpy(pts.view(np.int64)).cuda(); d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    d_out = torch.zeros(12, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctx.check(ctx.lib.trp_dev_msm_var(ctx.handle, d_pts.data_ptr(), d_sc.data_ptr(), n, 1, d_out.data_ptr()))
    ctx.sync()
    assert np.array_equal(O.jacobian_to_affine(O.VESTA, d_out.cpu().numpy().view(np.uint64).reshape(3, 4)), O.msm(O.VESTA, sc, pts))


class Transcript:
    """Deterministic stand-in for the Blake2b transcript: records what is written, challenges come from a seeded stream."""

    def __init__(self, seed, p):
        self.rng, self.p, self.log = random.Random(seed), p, []

    def write_point(self, P):
        self.log.append(("point", P))

    def write_scalar(self, s):
        self.log.append(("scalar", s))

    def squeeze_challenge_scalar(self):
        return self.rng.randrange(1, self.p)


@pytest.mark.parametrize("curve", [O.VESTA, O.PALLAS])
@pytest.mark.parametrize("k", [1, 4, 6])
def test_ipa_create_proof(pkg, ctxs, curve, k):
    ctx = ctxs[curve]
    field, F = FIELD_OF[curve]
    C = CURVE_OF[curve]
    rng = random.Random(100 + k)
    n = 1 << k
    G = (C.base.p - 1, 2)
    g = [C.mul(rng.randrange(1, F.p), G) for _ in range(n)]
    w, u = C.mul(rng.randrange(1, F.p), G), C.mul(rng.randrange(1, F.p), G)
    p_poly = [rng.randrange(F.p) for _ in range(n)]
    p_blind, x_3 = rng.randrange(F.p), rng.randrange(F.p)
    draws = [rng.randrange(F.p) for _ in range(n + 1 + 2 * k)]
    it1, it2 = iter(draws), iter(draws)
    t_want = Transcript(9, F.p)
    pm.ipa_create_proof(C, k, g, w, u, lambda: next(it1), t_want, p_poly, p_blind, x_3)
    params = pkg.ipa.IpaParams(ctx, k, np.stack([pt_limbs(curve, P) for P in g]), pt_limbs(curve, w), pt_limbs(curve, u))
    t_got = Transcript(9, F.p)
    pkg.ipa.create_proof(params, lambda: next(it2), t_got, mont(field, p_poly), p_blind, x_3)
    params.free()
    assert len(t_got.log) == len(t_want.log) == 1 + 2 * k + 2
    for (kind_g, val_g), (kind_w, val_w) in zip(t_got.log, t_want.log):
        assert kind_g == kind_w
        assert (pt_of(curve, val_g) if kind_g == "point" else val_g) == val_w
