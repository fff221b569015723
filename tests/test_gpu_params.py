"""GPU parity for SURVEY.md 8(f) row f3 -- hash-to-curve, the group FFT and Params::new -- through the C ABI, bit-exact against
oracle/params_model.py (itself pinned by pasta_curves' published vectors, tests/test_params_cpu.py)."""
import json
import os
import random

import numpy as np
import pytest

from util import O, pm

import params_model as M

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "params_kat.json")))
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


def pt_limbs(curve, P):
    if P is None:
        return np.zeros(8, dtype=np.uint64)
    return mont(O.BASE_FIELD[curve], [P[0], P[1]]).reshape(8)


def pts_limbs(curve, Ps):
    return np.stack([pt_limbs(curve, P) for P in Ps])


def pt_of(curve, limbs):
    if not np.asarray(limbs).any():
        return None
    x, y = O.limbs_to_ints(O.from_mont(O.BASE_FIELD[curve], np.asarray(limbs, dtype=np.uint64).reshape(2, 4)))
    return (x, y)


def test_published_known_answer_on_the_gpu(pkg, ctxs):
    kat = GOLD["published"]["pallas_hash_to_curve_kat"]
    x, y, z = (int(kat[k], 16) for k in ("jacobian_x", "jacobian_y", "jacobian_z"))
    p = pm.Fp.p
    zi = pow(z, -1, p)
    got = pkg.hash_to_curve(ctxs[O.PALLAS], kat["domain_prefix"])(kat["message"].encode())
    assert pt_of(O.PALLAS, got) == (x * zi * zi % p, y * zi ** 3 % p)


@pytest.mark.parametrize("curve", [O.VESTA, O.PALLAS])
@pytest.mark.parametrize("msg_len", [0, 1, 5, 17, 64, 75, 76, 200, 333])
def test_hash_to_curve_matches_oracle(pkg, ctxs, curve, msg_len):
    """message lengths straddle the BLAKE2b block boundary of the first hash (128 zero bytes + msg + 3 + len(DST'))"""
    C = CURVE_OF[curve]
    rnd = random.Random(msg_len)
    msgs = [bytes(rnd.randrange(256) for _ in range(msg_len)) for _ in range(1 if msg_len == 0 else 7)]
    for prefix in ("Halo2-Parameters", "z.cash:test", ""):
        got = pkg.hash_to_curve(ctxs[curve], prefix)(msgs)
        h = M.hash_to_curve(C, prefix)
        assert [pt_of(curve, g) for g in got] == [h(m) for m in msgs]


@pytest.mark.parametrize("curve", [O.VESTA, O.PALLAS])
def test_generators_against_fixture_and_oracle(pkg, ctxs, curve):
    import ctypes
    import torch
    ctx, C = ctxs[curve], CURVE_OF[curve]
    name = C.name
    n = 300
    out = torch.zeros((n, 8), dtype=torch.int64, device="cuda")
    zero = (ctypes.c_uint8 * 1)(0)
    ctx.check(ctx.lib.trp_dev_hash_to_curve(ctx.handle, b"Halo2-Parameters", zero, 1, 1, 0, n, out.data_ptr()))
    ctx.sync()
    got = [pt_of(curve, r) for r in out.cpu().numpy().view(np.uint64)]
    assert got[:16] == [tuple(int(v, 16) for v in e) for e in GOLD["curves"][name]["generators_0_15"]]
    assert got == M.params_generators(C, n)
    ctx.check(ctx.lib.trp_dev_hash_to_curve(ctx.handle, b"Halo2-Parameters", zero, 1, 1, 1 << 20, 2, out.data_ptr()))
    ctx.sync()
    got = [pt_of(curve, r) for r in out[:2].cpu().numpy().view(np.uint64)]
    assert got == [tuple(int(v, 16) for v in e) for e in GOLD["curves"][name]["generators_at_2^20"]]


@pytest.mark.parametrize("curve", [O.VESTA, O.PALLAS])
@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5])
def test_group_fft_matches_oracle(pkg, ctxs, curve, log_n):
    C = CURVE_OF[curve]
    Fs, sf = C.scalar, O.SCALAR_FIELD[curve]
    rnd = random.Random(100 + log_n)
    n = 1 << log_n
    pts = [C.mul(rnd.randrange(1, Fs.p), C.G) for _ in range(n)]
    if n >= 4:
        pts[1] = None                      # identity among the inputs
        pts[3] = pts[2]                    # equal inputs: the butterfly hits the doubling / cancelling branches
    omega = Fs.root_of_unity(log_n)
    for scale in (None, rnd.randrange(1, Fs.p)):
        got = pkg.best_fft_group(ctxs[curve], pts_limbs(curve, pts), mont(sf, [omega])[0], log_n,
                                 None if scale is None else mont(sf, [scale])[0])
        want = pm.best_fft_group(C, pts, omega, log_n)
        if scale is not None:
            want = [C.mul(scale, P) for P in want]
        assert [pt_of(curve, g) for g in got] == want


@pytest.mark.parametrize("curve", [O.VESTA, O.PALLAS])
def test_params_new_small_matches_oracle(pkg, ctxs, curve):
    C = CURVE_OF[curve]
    for k in (0, 1, 3, 4):
        prm = pkg.Params.new(ctxs[curve], k)
        want = M.params_new(C, k)
        assert [pt_of(curve, g) for g in prm.g_points] == want["g"]
        assert [pt_of(curve, g) for g in prm.g_lagrange_points] == want["g_lagrange"]
        assert pt_of(curve, prm.w) == want["w"] and pt_of(curve, prm.u) == want["u"]
    fx = GOLD["curves"][C.name]["params_k3"]
    prm = pkg.Params.new(ctxs[curve], 3)
    assert [pt_of(curve, g) for g in prm.g_lagrange_points] == [tuple(int(v, 16) for v in e) for e in fx["g_lagrange"]]


@pytest.mark.parametrize("curve", [O.VESTA, O.PALLAS])
def test_params_new_k12_commit_lagrange_equals_commit_of_coefficients(pkg, ctxs, curve):
    """size-independent property of the whole of Params::new: commit_lagrange(v, r) == commit(lagrange_to_coeff(v), r), and
    the group iFFT followed by the forward group FFT is the identity"""
    ctx, C = ctxs[curve], CURVE_OF[curve]
    sf = O.SCALAR_FIELD[curve]
    k = 12
    prm = pkg.Params.new(ctx, k)
    v = O.random_field_mont(sf, 1 << k, 77)
    blind = O.random_field_mont(sf, 1, 78)[0]
    dom = pkg.EvaluationDomain(ctx, 3, k)
    a = prm.commit_lagrange(v, blind)
    b = prm.commit(dom.lagrange_to_coeff(v), blind)
    assert np.array_equal(a, b) and a[2].any()
    n_mont = mont(sf, [1 << k])[0]
    back = pkg.best_fft_group(ctx, prm.g_lagrange_points, dom.omega, k)
    assert np.array_equal(back, prm.g_points)
    del n_mont
    # spot-check generators against the oracle
    want = M.params_generators(C, 4, start=(1 << k) - 4)
    assert [pt_of(curve, g) for g in prm.g_points[-4:]] == want


def test_hash_to_curve_argument_errors(pkg, ctxs):
    ctx = ctxs[O.VESTA]
    with pytest.raises(pkg.TrpError):
        pkg.hash_to_curve(ctx, "x" * 250)(b"m")
    with pytest.raises(ValueError):
        pkg.hash_to_curve(ctx, "p")([b"a", b"bc"])
    with pytest.raises(ValueError):
        pkg.Params.new(ctx, 32)


def test_random_field_is_blake2b_from_u512():
    """trp_dev_random_field: out[i] = from_u512(BLAKE2b-512(key ++ u64_le(first + i))) in Montgomery form, both fields"""
    import hashlib
    import __graft_entry__ as ge
    import torch
    pkg = ge.load_package()
    for curve, field in ((pkg.VESTA, O.FP), (pkg.PALLAS, O.FQ)):
        ctx = pkg.Context(0, curve)
        ctx.bind_torch_stream()
        key, first, n = bytes(range(7, 39)), (1 << 40) + 5, 1000
        out = torch.zeros((n, 4), dtype=torch.int64, device="cuda")
        ctx.check(ctx.lib.trp_dev_random_field(ctx.handle, 0, key, first, n, out.data_ptr()))
        got = O.limbs_to_ints(O.from_mont(field, out.cpu().numpy().view(np.uint64)))
        p = O.MODULUS[field]
        want = [int.from_bytes(hashlib.blake2b(key + (first + i).to_bytes(8, "little"), digest_size=64).digest(), "little") % p for i in range(n)]
        assert got == want
