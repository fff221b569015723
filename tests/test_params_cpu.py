"""CPU tests of the row-f3 oracle (oracle/params_model.py: pasta_curves' hash-to-curve and halo2's Params::new) against the
PUBLISHED vectors of pasta_curves 0.4.1 -- the one part of the path whose parity is pinned by the dependency's own
known-answer data -- and against the frozen fixture tests/golden/params_kat.json."""
import json
import os
import random

import pytest

from util import pm

import params_model as M

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "params_kat.json")))
CURVES = {"pallas": pm.Pallas, "vesta": pm.Vesta}


def _pt(v):
    return None if v is None else (int(v[0], 16), int(v[1], 16))


def _iso_mul(F, a, k, P):
    R = None
    while k:
        if k & 1:
            R = M._iso_add(F, a, R, P)
        P = M._iso_add(F, a, P, P)
        k >>= 1
    return R


@pytest.mark.parametrize("name", ["pallas", "vesta"])
def test_iso_curve_has_the_curve_order(name):
    """y^2 = x^3 + a x + 1265 is isogenous to the target curve, hence has the same number of points"""
    C = CURVES[name]
    F, a = C.base, M.ISO_A[name]
    rnd = random.Random(3)
    for _ in range(3):
        while True:
            x = rnd.randrange(F.p)
            y = F.sqrt((x ** 3 + a * x + M.ISO_B) % F.p)
            if y is not None:
                break
        assert _iso_mul(F, a, C.scalar.p, (x, y)) is None


def test_derived_isogeny_equals_published_pallas_constants():
    want = [int(v, 16) for v in GOLD["published"]["pallas_isogeny_constants"]]
    assert M.derive_isogeny_constants(pm.Fp, M.ISO_A["pallas"]) == want


@pytest.mark.parametrize("name", ["pallas", "vesta"])
def test_isogeny_is_a_homomorphism_onto_the_curve(name):
    C = CURVES[name]
    F, a, c = C.base, M.ISO_A[name], M.isogeny_constants(C)
    assert [hex(v) for v in c] == GOLD["curves"][name]["isogeny_constants"]
    rnd = random.Random(4)
    pts = []
    while len(pts) < 3:
        x = rnd.randrange(F.p)
        y = F.sqrt((x ** 3 + a * x + M.ISO_B) % F.p)
        if y is not None:
            pts.append((x, y))
    P, Q, _ = pts
    for X in pts:
        assert C.on_curve(M.iso_map(F, c, X))
    assert M.iso_map(F, c, M._iso_add(F, a, P, Q)) == C.add(M.iso_map(F, c, P), M.iso_map(F, c, Q))
    assert M.iso_map(F, c, M._iso_add(F, a, P, P)) == C.double(M.iso_map(F, c, P))


def test_published_hash_to_curve_known_answer():
    """pasta_curves' own test vector: hash_to_curve("z.cash:test")(b"Trans rights now!") on Pallas (Jacobian coordinates)"""
    kat = GOLD["published"]["pallas_hash_to_curve_kat"]
    x, y, z = (int(kat[k], 16) for k in ("jacobian_x", "jacobian_y", "jacobian_z"))
    p = pm.Fp.p
    zi = pow(z, -1, p)
    want = (x * zi * zi % p, y * zi ** 3 % p)
    got = M.hash_to_curve(pm.Pallas, kat["domain_prefix"])(kat["message"].encode())
    assert got == want and pm.Pallas.on_curve(got)


@pytest.mark.parametrize("name", ["pallas", "vesta"])
def test_fixture_matches_oracle(name):
    C, g = CURVES[name], GOLD["curves"][name]
    h = M.hash_to_curve(C, "Halo2-Parameters")
    for e in g["hash_to_field"]:
        assert [hex(u) for u in M.hash_to_field(C.base, name, "Halo2-Parameters", bytes.fromhex(e["message"]))] == e["u"]
    for e in g["hash_to_curve"]:
        assert h(bytes.fromhex(e["message"])) == _pt(e["point"])
    gens = M.params_generators(C, 16)
    assert gens == [_pt(v) for v in g["generators_0_15"]]
    assert all(C.on_curve(P) and P is not None for P in gens) and len(set(gens)) == 16


@pytest.mark.parametrize("name", ["pallas", "vesta"])
def test_swu_edge_cases(name):
    """u = 0 (ta = 0 branch) and sign rule: the image is on the iso-curve and y has u's parity"""
    C = CURVES[name]
    F, a = C.base, M.ISO_A[name]
    for u in (0, 1, 2, F.p - 1, 12345678901234567890):
        x, y = M.map_to_curve_simple_swu(F, a, M.ISO_B, M.SWU_Z, u)
        assert (y * y - (x ** 3 + a * x + M.ISO_B)) % F.p == 0
        assert (y & 1) == (u & 1)


@pytest.mark.parametrize("name", ["pallas", "vesta"])
def test_g_lagrange_commits_like_g(name):
    """the defining property of Params::new's group iFFT: commit_lagrange(v) == commit(lagrange_to_coeff(v))"""
    C = CURVES[name]
    k = 3
    prm = M.params_new(C, k)
    assert prm["g"] == [_pt(v) for v in GOLD["curves"][name]["params_k3"]["g"]]
    assert prm["g_lagrange"] == [_pt(v) for v in GOLD["curves"][name]["params_k3"]["g_lagrange"]]
    assert prm["w"] == _pt(GOLD["curves"][name]["params_k3"]["w"]) and prm["u"] == _pt(GOLD["curves"][name]["params_k3"]["u"])
    Fs = C.scalar
    rnd = random.Random(9)
    v = [rnd.randrange(Fs.p) for _ in range(1 << k)]
    dom = pm.EvaluationDomain(Fs, 3, k)
    assert C.naive_msm(v, prm["g_lagrange"]) == C.naive_msm(dom.lagrange_to_coeff(v), prm["g"])


# ---- host build of the device code (csrc/h2c.cuh) against the oracle: the byte / limb logic of the CUDA kernel on the CPU box ------
@pytest.fixture(scope="module")
def h2c_shim():
    import ctypes
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    so = os.path.join(here, "_h2c_host_shim.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", os.path.join(here, "h2c_host_shim.cpp"), "-o", so])
    return ctypes.CDLL(so)


def _shim_hash(shim, C, prefix, msgs):
    import ctypes
    import numpy as np
    n, ml = len(msgs), len(msgs[0])
    buf = np.frombuffer(b"".join(msgs) + b"\0", dtype=np.uint8).copy()
    out = np.zeros((n, 16), dtype=np.uint32)
    rc = shim.h2ch_hash(int(C.name == "pallas"), prefix.encode(), buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(ml),
                        ctypes.c_size_t(n), out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    F = C.base
    val = lambda r: sum(int(v) << (32 * i) for i, v in enumerate(r))
    res = []
    for r in out:
        x, y = F.from_mont(val(r[:8])), F.from_mont(val(r[8:]))
        res.append(None if (x, y) == (0, 0) else (x, y))
    return res


@pytest.mark.parametrize("name", ["pallas", "vesta"])
def test_device_hash_to_curve_code_on_host(h2c_shim, name):
    C = CURVES[name]
    rnd = random.Random(11)
    kat = GOLD["published"]["pallas_hash_to_curve_kat"]
    if name == "pallas":
        got = _shim_hash(h2c_shim, C, kat["domain_prefix"], [kat["message"].encode()])
        assert got == [M.hash_to_curve(C, kat["domain_prefix"])(kat["message"].encode())]
    for ml in (0, 1, 5, 75, 76, 77, 130, 300):        # around the BLAKE2b block boundary of the first hash
        msgs = [bytes(rnd.randrange(256) for _ in range(ml)) for _ in range(1 if ml == 0 else 12)]
        h = M.hash_to_curve(C, "Halo2-Parameters")
        assert _shim_hash(h2c_shim, C, "Halo2-Parameters", msgs) == [h(m) for m in msgs]
    gens = [bytes([0]) + i.to_bytes(4, "little") for i in range(64)]
    assert _shim_hash(h2c_shim, C, "Halo2-Parameters", gens) == M.params_generators(C, 64)
