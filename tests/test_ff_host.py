"""CPU unit tests of the limb-level logic in csrc/ff.cuh / ec.cuh (the plain-C twins of the PTX carry chains),
checked against the Python big-int model.  The same algorithms run as PTX on the GPU (tests/test_gpu_*.py)."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

import pasta_model as pm
import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_ff_host_shim.so")


@pytest.fixture(scope="module")
def shim():
    src = os.path.join(HERE, "ff_host_shim.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", src, "-o", SO])
    return ctypes.CDLL(SO)


def _run(shim, field, op, a, b):
    A = O.ints_to_limbs(a); B = O.ints_to_limbs(b); R = np.zeros_like(A)
    shim.ffh_op(field, op, A.ctypes.data_as(ctypes.c_void_p), B.ctypes.data_as(ctypes.c_void_p),
                R.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(a)))
    return O.limbs_to_ints(R)


@pytest.mark.parametrize("fid,F", [(0, pm.Fp), (1, pm.Fq)])
def test_field_ops_match_bigint(shim, fid, F):
    rnd = random.Random(7 + fid)
    p = F.p
    edge = [0, 1, 2, p - 1, p - 2, F.R, F.R2, (1 << 254) - 1, 1 << 254, (1 << 32) - 1, 1 << 32, p >> 1]
    a = edge + [rnd.randrange(p) for _ in range(500)]
    b = list(reversed(edge)) + [rnd.randrange(p) for _ in range(500)]
    a += edge; b += edge
    Ri = F.Rinv
    assert _run(shim, fid, 0, a, b) == [(x + y) % p for x, y in zip(a, b)]
    assert _run(shim, fid, 1, a, b) == [(x - y) % p for x, y in zip(a, b)]
    assert _run(shim, fid, 2, a, b) == [x * y * Ri % p for x, y in zip(a, b)]        # Montgomery product
    assert _run(shim, fid, 4, a, b) == [x * x * Ri % p for x in a]
    assert _run(shim, fid, 5, a, b) == [x * Ri % p for x in a]
    assert _run(shim, fid, 6, a, b) == [x * F.R % p for x in a]
    assert _run(shim, fid, 7, a, b) == [(-x) % p for x in a]
    inv = _run(shim, fid, 3, a[:40], b[:40])
    # mont inverse: (aR)^-1 * R^2 * R^-1... fe_inv(x) = x^(p-2) in Montgomery arithmetic => result r with r*x*Ri = R
    for x, r in zip(a[:40], inv):
        if x == 0:
            assert r == 0
        else:
            assert r * x * Ri % p == F.R


@pytest.mark.parametrize("cid,C", [(0, pm.Pallas), (1, pm.Vesta)])
def test_xyzz_ops_match_bigint(shim, cid, C):
    rnd = random.Random(11 + cid)
    F = C.base
    bf = O.BASE_FIELD[cid]

    def aff(P):
        return [0, 0] if P is None else [F.to_mont(P[0]), F.to_mont(P[1])]

    def run(points, negs, q=None, dbls=0):
        pts = O.ints_to_limbs([c for P in points for c in aff(P)])
        ng = np.array(negs, dtype=np.int32)
        out = np.zeros((2, 4), dtype=np.uint64)
        qarr = None
        if q is not None:
            # q as XYZZ with zz = z^2, zzz = z^3 for a random z
            z = rnd.randrange(1, F.p)
            if q is None:
                pass
            X = q[0] * z * z % F.p; Y = q[1] * pow(z, 3, F.p) % F.p
            qarr = O.ints_to_limbs([F.to_mont(X), F.to_mont(Y), F.to_mont(z * z % F.p), F.to_mont(pow(z, 3, F.p))])
        shim.ffh_ecop(bf, pts.ctypes.data_as(ctypes.c_void_p), ng.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(points)),
                      qarr.ctypes.data_as(ctypes.c_void_p) if qarr is not None else None, dbls, out.ctypes.data_as(ctypes.c_void_p))
        x, y = [F.from_mont(v) for v in O.limbs_to_ints(out)]
        return None if (x == 0 and y == 0) else (x, y)

    P = [C.mul(rnd.randrange(C.scalar.p), C.G) for _ in range(6)]
    exp = None
    for p_ in P:
        exp = C.add(exp, p_)
    assert run(P, [0] * 6) == exp
    assert run(P, [1] * 6) == C.neg(exp)
    # doubling path (P + P), cancellation (P - P), identity inputs
    assert run([P[0], P[0]], [0, 0]) == C.double(P[0])
    assert run([P[0], P[0], P[1]], [0, 1, 0]) == P[1]
    assert run([P[0], P[0]], [0, 1]) is None
    assert run([None, P[2], None], [0, 0, 0]) == P[2]
    # full add incl. special cases, and repeated doubling
    assert run(P[:3], [0, 0, 0], q=P[4]) == C.add(C.add(C.add(P[0], P[1]), P[2]), P[4])
    assert run([P[0]], [0], q=P[0]) == C.double(P[0])
    assert run([P[0]], [0], q=C.neg(P[0])) is None
    assert run([], [], q=P[3]) == P[3]
    assert run([P[0]], [0], q=P[1], dbls=5) == C.mul(32, C.add(P[0], P[1]))


@pytest.mark.parametrize("cid,C", [(0, pm.Pallas), (1, pm.Vesta)])
@pytest.mark.parametrize("threads,log_m,log_chunk", [(8, 2, 2), (4, 0, 3), (1, 3, 1), (16, 1, 0)])
def test_two_level_bucket_sum_matches_bigint(shim, cid, C, threads, log_m, log_chunk):
    """csrc/bucket_reduce.cuh: sum_b (b + 1) B_b through chunk sums, per-thread running sums over the chunk totals and the
    (sum, weighted sum) tree across threads, with empty buckets, repeated points (the doubling case) and cancelling pairs"""
    rnd = random.Random(31 + cid + 7 * threads)
    F = C.base
    bf = O.BASE_FIELD[cid]
    nb = threads << (log_m + log_chunk)
    P = [C.mul(rnd.randrange(C.scalar.p), C.G) for _ in range(12)]
    buckets = [rnd.choice(P + [None, None, None, C.neg(P[0]), C.neg(P[1])]) for _ in range(nb)]
    aff = lambda Q: [0, 0] if Q is None else [F.to_mont(Q[0]), F.to_mont(Q[1])]
    arr = O.ints_to_limbs([c for Q in buckets for c in aff(Q)])
    out = np.zeros((2, 4), dtype=np.uint64)
    shim.ffh_bucket_reduce2(bf, arr.ctypes.data_as(ctypes.c_void_p), threads, log_m, log_chunk, out.ctypes.data_as(ctypes.c_void_p))
    vals = [F.from_mont(v) for v in O.limbs_to_ints(out)]
    got = None if vals == [0, 0] else (vals[0], vals[1])
    want = None
    for b, Q in enumerate(buckets):
        want = C.add(want, C.mul(b + 1, Q) if Q is not None else None)
    assert got == want
