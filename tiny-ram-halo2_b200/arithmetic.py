"""Mirror of halo2_proofs::arithmetic::{best_multiexp, best_fft} (halo2_proofs 0.2.0, un-vendored dependency of the
reference, Cargo.lock:619-621) on top of the C ABI.  Same argument meaning and error behaviour: halo2 panics on
``assert_eq!(coeffs.len(), bases.len())`` / ``assert_eq!(a.len(), 1 << log_n)``; here those raise ``TrpError``/``ValueError``.

Arrays are numpy uint64 (..., 4) Montgomery limbs -- the in-memory form of pasta_curves' Fp / Fq."""
from __future__ import annotations

import ctypes

import numpy as np

from ._lib import Context, as_u64, ptr


class Bases:
    """Device-resident MSM bases (Params.g / g_lagrange ++ [w]): uploaded once, reused by every commitment."""

    def __init__(self, ctx: Context, affine_xy, flags: int = 0):
        self.ctx = ctx
        xy = as_u64(affine_xy)
        if xy.size % 8:
            raise ValueError("bases must be n x 8 uint64 (x[4], y[4])")
        self.n = xy.size // 8
        h = ctypes.c_void_p()
        ctx.check(ctx.lib.trp_bases_load_ex(ctx.handle, ptr(xy), self.n, flags, ctypes.byref(h)))
        self.handle = h

    def describe(self):
        out = (ctypes.c_uint * 3)()
        self.ctx.check(self.ctx.lib.trp_bases_describe(self.handle, out))
        return {"c": out[0], "windows": out[1], "precomputed": bool(out[2])}

    def free(self):
        if getattr(self, "handle", None) and getattr(self.ctx, "handle", None):
            self.ctx.lib.trp_bases_free(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def best_multiexp(ctx: Context, coeffs, bases: Bases):
    """best_multiexp(coeffs, bases) -> C::Curve: Jacobian (x, y, z) Montgomery limbs, shape (3, 4); identity <=> z = 0.
    ``coeffs`` may be (n, 4) for one MSM or (m, n, 4) for m MSMs over the same bases (returns (m, 3, 4))."""
    sc = as_u64(coeffs)
    if sc.ndim < 2 or sc.shape[-1] != 4:
        raise ValueError("coeffs must be (n, 4) or (m, n, 4) uint64")
    single = sc.ndim == 2
    m = 1 if single else sc.shape[0]
    n = sc.shape[-2]
    if n > bases.n:
        raise ValueError(f"coeffs.len() = {n} exceeds bases.len() = {bases.n}")
    out = np.zeros((m, 3, 4), dtype=np.uint64)
    ctx.check(ctx.lib.trp_msm_batch(ctx.handle, bases.handle, ptr(sc), n, m, ptr(out)))
    return out[0] if single else out


def best_fft(ctx: Context, a, omega, log_n: int):
    """best_fft(a, omega, log_n): natural order in/out radix-2 NTT; ``a`` is (2^log_n, 4) or (batch, 2^log_n, 4).
    Returns the transformed array (halo2 transforms in place; numpy callers get a new array)."""
    arr = as_u64(a, copy=True)
    if arr.shape[-1] != 4 or arr.shape[-2] != (1 << log_n):
        raise ValueError(f"a.len() = {arr.shape[-2] if arr.ndim > 1 else arr.size} != 1 << log_n = {1 << log_n}")
    batch = arr.size // (4 << log_n)
    om = as_u64(omega).reshape(4)
    ctx.check(ctx.lib.trp_ntt(ctx.handle, ptr(arr), batch, log_n, ptr(om)))
    return arr


def best_fft_group(ctx: Context, points, omega, log_n: int, scale=None):
    """best_fft over curve points (the `Group` instance Params::new uses): ``points`` is (2^log_n, 8) normalised affine
    Montgomery limbs (identity = all zero); out[k] = sum_j [omega^(jk)] points[j], every output times ``scale`` when given.
    Returns a new array."""
    arr = as_u64(points, copy=True)
    if arr.size != 8 << log_n:
        raise ValueError(f"a.len() = {arr.size // 8} != 1 << log_n = {1 << log_n}")
    om = as_u64(omega).reshape(4)
    sc = None if scale is None else as_u64(scale).reshape(4)
    ctx.check(ctx.lib.trp_group_fft(ctx.handle, ptr(arr), log_n, ptr(om), ptr(sc)))
    return arr.reshape(-1, 8)


def hash_to_curve(ctx: Context, domain_prefix: str):
    """CurveExt::hash_to_curve(domain_prefix) of the ctx's curve -> closure taking one message (bytes) or a list of
    equal-length messages and returning (8,) / (n, 8) affine Montgomery limbs."""
    prefix = domain_prefix.encode()

    def hasher(message):
        single = isinstance(message, (bytes, bytearray))
        msgs = [bytes(message)] if single else [bytes(m) for m in message]
        if len({len(m) for m in msgs}) > 1:
            raise ValueError("messages of one call must have equal length")
        buf = np.frombuffer(b"".join(msgs), dtype=np.uint8) if msgs and len(msgs[0]) else np.zeros(1, dtype=np.uint8)
        out = np.zeros((len(msgs), 8), dtype=np.uint64)
        ctx.check(ctx.lib.trp_hash_to_curve(ctx.handle, prefix, ptr(buf), len(msgs[0]) if msgs else 0, len(msgs), ptr(out)))
        return out[0] if single else out

    return hasher
