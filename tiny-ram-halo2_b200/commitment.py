"""Mirror of halo2_proofs::poly::commitment::Params::{commit, commit_lagrange} (poly/commitment.rs, halo2_proofs 0.2.0):
both append ``blind * w`` to the polynomial and call best_multiexp over ``g ++ [w]`` / ``g_lagrange ++ [w]``.
``Params.new(ctx, k)`` mirrors Params::new (reached from src/test_utils.rs:21,89): hash-to-curve
generators and the group iFFT to the Lagrange basis run on the device (csrc/params.cu)."""
from __future__ import annotations

import numpy as np

from ._lib import Context, as_u64, ptr
from .arithmetic import Bases, best_multiexp


class Params:
    def __init__(self, ctx: Context, k: int, g, g_lagrange, w, flags: int = 0, u=None):
        self.ctx, self.k, self.n = ctx, k, 1 << k
        g = as_u64(g).reshape(-1, 8); gl = as_u64(g_lagrange).reshape(-1, 8); w = as_u64(w).reshape(1, 8)
        if len(g) != self.n or len(gl) != self.n:
            raise ValueError("g and g_lagrange must hold 2^k points")
        self.g_points, self.g_lagrange_points, self.w, self.u = g, gl, w[0], (None if u is None else as_u64(u).reshape(8))
        self.g = Bases(ctx, np.concatenate([g, w]), flags)
        self.g_lagrange = Bases(ctx, np.concatenate([gl, w]), flags)

    @classmethod
    def new(cls, ctx: Context, k: int, flags: int = 0):
        """Params::new(k): g[i] = hash_to_curve("Halo2-Parameters")([0] ++ u32_le(i)), g_lagrange by the group iFFT,
        w = hash([1]), u = hash([2]).  halo2 asserts k < 32."""
        if not 0 <= k < 32:
            raise ValueError("k must be below 32")
        n = 1 << k
        g = np.zeros((n, 8), dtype=np.uint64); gl = np.zeros((n, 8), dtype=np.uint64)
        w = np.zeros(8, dtype=np.uint64); u = np.zeros(8, dtype=np.uint64)
        ctx.check(ctx.lib.trp_params_new(ctx.handle, k, ptr(g), ptr(gl), ptr(w), ptr(u)))
        return cls(ctx, k, g, gl, w, flags, u=u)

    def _commit(self, bases, poly, blind):
        poly = as_u64(poly)
        if poly.shape[-2] != self.n:
            raise ValueError("polynomial length must be 2^k")
        blind = as_u64(blind)
        if poly.ndim == 2:
            sc = np.concatenate([poly, blind.reshape(1, 4)])
        else:
            sc = np.concatenate([poly, blind.reshape(poly.shape[0], 1, 4)], axis=1)
        return best_multiexp(self.ctx, sc, bases)

    def commit(self, poly, blind):
        """Params::commit(poly in coefficient basis, Blind) -> C::Curve"""
        return self._commit(self.g, poly, blind)

    def commit_lagrange(self, poly, blind):
        """Params::commit_lagrange(poly in Lagrange basis, Blind) -> C::Curve (an MSM of n + 1 points)"""
        return self._commit(self.g_lagrange, poly, blind)
