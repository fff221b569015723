"""ctypes loader for the C++ CPU oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
The product package (tiny-ram-halo2_b200/) never imports this module.

Arrays are numpy uint64 of shape (..., 4) holding Montgomery-form little-endian limbs, i.e. exactly
what pasta_curves' Fp/Fq hold in memory and what the C ABI in include/tr_prover.h takes.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

FP, FQ = 0, 1
PALLAS, VESTA = 0, 1
SCALAR_FIELD = {PALLAS: FQ, VESTA: FP}
BASE_FIELD = {PALLAS: FP, VESTA: FQ}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_hw_threads.restype = ctypes.c_int
        _lib.orc_domain_info.restype = ctypes.c_uint
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _u64(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a


def hw_threads() -> int:
    return lib().orc_hw_threads()


def to_mont(field, canon):
    canon = _u64(canon); out = np.empty_like(canon)
    lib().orc_to_mont(field, _p(canon), _p(out), ctypes.c_size_t(canon.size // 4))
    return out


def from_mont(field, mont):
    mont = _u64(mont); out = np.empty_like(mont)
    lib().orc_from_mont(field, _p(mont), _p(out), ctypes.c_size_t(mont.size // 4))
    return out


def field_op(field, op, a, b=None):
    code = {"add": 0, "sub": 1, "mul": 2, "inv": 3, "sqr": 4}[op]
    a = _u64(a); out = np.empty_like(a)
    bp = _p(_u64(b)) if b is not None else None
    lib().orc_field_op(field, code, _p(a), bp, _p(out), ctypes.c_size_t(a.size // 4))
    return out


def fft(field, a, log_n, omega_mont, threads=None):
    """best_fft on one vector (returns a new array)."""
    a = _u64(a).copy()
    assert a.size == 4 << log_n
    om = _u64(omega_mont)
    lib().orc_fft(field, _p(a), ctypes.c_uint(log_n), _p(om), threads or hw_threads())
    return a


def msm(curve, scalars, bases, threads=None):
    """best_multiexp -> affine (x, y) Montgomery, shape (8,), identity = zeros."""
    scalars = _u64(scalars); bases = _u64(bases)
    n = scalars.size // 4
    assert bases.size == 8 * n
    out = np.zeros(8, dtype=np.uint64)
    lib().orc_msm(curve, _p(scalars), _p(bases), ctypes.c_size_t(n), threads or hw_threads(), _p(out))
    return out


def points_progression(curve, p0, d, n, threads=None):
    p0 = _u64(p0); d = _u64(d)
    out = np.empty((n, 8), dtype=np.uint64)
    lib().orc_points_progression(curve, _p(p0), _p(d), ctypes.c_size_t(n), threads or hw_threads(), _p(out))
    return out


def point_mul(curve, k_canon, p):
    k = _u64(k_canon); p = _u64(p); out = np.zeros(8, dtype=np.uint64)
    lib().orc_point_mul(curve, _p(k), _p(p), _p(out))
    return out


def point_add(curve, a, b):
    a = _u64(a); b = _u64(b); out = np.zeros(8, dtype=np.uint64)
    lib().orc_point_add(curve, _p(a), _p(b), _p(out))
    return out


def jacobian_to_affine(curve, jac):
    """(..., 3, 4) Jacobian Montgomery limbs -> (..., 8) affine; identity -> zeros."""
    jac = _u64(jac); n = jac.size // 12
    out = np.zeros((n, 8), dtype=np.uint64)
    lib().orc_jacobian_to_affine(curve, _p(jac), ctypes.c_size_t(n), _p(out))
    return out[0] if jac.ndim == 2 else out.reshape(jac.shape[:-2] + (8,))


def point_compress(curve, pts):
    pts = _u64(pts); n = pts.size // 8
    out = np.zeros((n, 32), dtype=np.uint8)
    lib().orc_point_compress(curve, _p(pts), ctypes.c_size_t(n), _p(out))
    return out


def lagrange_to_coeff(field, j, k, cols, threads=None):
    cols = _u64(cols).copy()
    batch = cols.size // (4 << k)
    lib().orc_lagrange_to_coeff(field, ctypes.c_uint(j), ctypes.c_uint(k), _p(cols), ctypes.c_size_t(batch), threads or hw_threads())
    return cols


def domain_info(field, j, k):
    om = np.zeros(4, dtype=np.uint64); eom = np.zeros(4, dtype=np.uint64)
    ek = lib().orc_domain_info(field, ctypes.c_uint(j), ctypes.c_uint(k), _p(om), _p(eom))
    return ek, om, eom


def coeff_to_extended(field, j, k, coeff, threads=None):
    coeff = _u64(coeff)
    batch = coeff.size // (4 << k)
    ek, _, _ = domain_info(field, j, k)
    ext = np.empty((batch, 1 << ek, 4), dtype=np.uint64)
    lib().orc_coeff_to_extended(field, ctypes.c_uint(j), ctypes.c_uint(k), _p(coeff), _p(ext), ctypes.c_size_t(batch), threads or hw_threads())
    return ext


def extended_to_coeff(field, j, k, ext, divide=False, threads=None):
    ext = _u64(ext).copy()
    out = np.empty(((j - 1) << k, 4), dtype=np.uint64)
    lib().orc_extended_to_coeff(field, ctypes.c_uint(j), ctypes.c_uint(k), _p(ext), _p(out), int(bool(divide)), threads or hw_threads())
    return out


# ---- helpers shared by tests / bench (numpy <-> python ints) ------------------------------------
def ints_to_limbs(vals):
    """list of python ints -> (n, 4) uint64 array."""
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        for j in range(4):
            out[i, j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def limbs_to_ints(arr):
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 4)
    return [int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192 for r in arr]


MODULUS = {
    FP: 0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001,
    FQ: 0x40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001,
}


def random_field_mont(field, n, seed):
    """n uniform elements of the field, Montgomery form, PCG64(seed), rejection sampled."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p = MODULUS[field]
    top = np.uint64(p >> 192)
    out = np.empty((n, 4), dtype=np.uint64)
    filled = 0
    while filled < n:
        m = n - filled
        cand = rng.integers(0, 1 << 64, size=(m + 16, 4), dtype=np.uint64)
        cand[:, 3] &= np.uint64((1 << 63) - 1)          # 255-bit candidates
        ok = cand[:, 3] < top                            # strictly below the top limb => < p (rejects ~1/2)
        cand = cand[ok][:m]
        out[filled:filled + len(cand)] = cand
        filled += len(cand)
    # a uniform canonical value re-read as Montgomery form is still uniform; no conversion needed
    return out


# ---- pieces of the sampled create_proof baseline (bench.py's CPU legs) -------------------------------------------------------
def quotient_vm(field, code, n_regs, consts_mont, cols, col_rows, xs, rows, threads=None):
    """the straight-line quotient program (include/tr_prover.h instruction set) interpreted on the CPU for `rows` rows;
    cols: list of (col_rows, 4) Montgomery arrays indexed modulo col_rows; xs: (col_rows, 4) values COSETX reads"""
    code = np.ascontiguousarray(code, dtype=np.uint32).reshape(-1, 4)
    consts = _u64(consts_mont)
    cols = [_u64(c) for c in cols]
    assert col_rows & (col_rows - 1) == 0 and all(c.size == 4 * col_rows for c in cols)
    tab = (ctypes.c_void_p * max(len(cols), 1))(*[c.ctypes.data for c in cols])
    xs = _u64(xs)
    out = np.zeros((rows, 4), dtype=np.uint64)
    lib().orc_quotient_vm(field, code.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(code)), ctypes.c_uint(n_regs), _p(consts), tab,
                          ctypes.c_size_t(col_rows), _p(xs), ctypes.c_size_t(rows), _p(out), threads or hw_threads())
    return out


def eval_polynomial(field, coeffs, x_mont, threads=None):
    coeffs = _u64(coeffs)
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_eval_polynomial(field, _p(coeffs), ctypes.c_size_t(coeffs.size // 4), _p(_u64(x_mont)), threads or hw_threads(), _p(out))
    return out


def generator_collapse(curve, g_affine, u_canon_limbs, threads=None):
    """one round of parallel_generator_collapse: returns g_lo + [u] g_hi (half the points)"""
    g = _u64(g_affine).copy()
    half = g.size // 16
    lib().orc_generator_collapse(curve, _p(g), ctypes.c_size_t(half), _p(_u64(u_canon_limbs)), threads or hw_threads())
    return g.reshape(-1, 8)[:half]
