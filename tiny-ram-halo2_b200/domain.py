"""Mirror of halo2_proofs::poly::EvaluationDomain (poly/domain.rs, halo2_proofs 0.2.0) on top of the C ABI."""
from __future__ import annotations

import ctypes

import numpy as np

from ._lib import Context, as_u64, ptr


class EvaluationDomain:
    """EvaluationDomain::new(j, k): n = 2^k rows, extended domain 2^extended_k >= n * (j - 1), coset generator zeta."""

    def __init__(self, ctx: Context, j: int, k: int):
        self.ctx, self.j, self.k = ctx, j, k
        h = ctypes.c_void_p()
        ctx.check(ctx.lib.trp_domain_create(ctx.handle, k, j, ctypes.byref(h)))
        self.handle = h
        self.extended_k = int(ctx.lib.trp_domain_extended_k(h))
        consts = np.zeros((4, 4), dtype=np.uint64)
        ctx.check(ctx.lib.trp_domain_constants(h, ptr(consts)))
        self.omega, self.extended_omega, self.g_coset, self.g_coset_inv = consts
        self.n = 1 << k

    def extended_len(self):
        return 1 << self.extended_k

    def _cols(self, a, length):
        arr = as_u64(a, copy=True)
        if arr.shape[-1] != 4 or arr.shape[-2] != length:
            raise ValueError(f"expected vectors of {length} field elements")
        return arr, arr.size // (4 * length)

    def lagrange_to_coeff(self, a):
        arr, batch = self._cols(a, self.n)
        self.ctx.check(self.ctx.lib.trp_lagrange_to_coeff(self.handle, ptr(arr), batch))
        return arr

    def coeff_to_lagrange(self, a):
        arr, batch = self._cols(a, self.n)
        self.ctx.check(self.ctx.lib.trp_coeff_to_lagrange(self.handle, ptr(arr), batch))
        return arr

    def coeff_to_extended(self, a):
        arr, batch = self._cols(a, self.n)
        out = np.empty(arr.shape[:-2] + (self.extended_len(), 4), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.trp_coeff_to_extended(self.handle, ptr(arr), ptr(out), batch))
        return out

    def extended_to_coeff(self, a, divide_by_vanishing_poly: bool = False):
        """extended_to_coeff(a) -> n*(j-1) coefficients; with divide_by_vanishing_poly=True the input is first
        multiplied by 1/(X^n - 1) on the coset (EvaluationDomain::divide_by_vanishing_poly)."""
        arr, batch = self._cols(a, self.extended_len())
        if batch != 1:
            raise ValueError("extended_to_coeff takes one extended polynomial")
        out = np.empty((self.n * (self.j - 1), 4), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.trp_extended_to_coeff(self.handle, ptr(arr), ptr(out), int(divide_by_vanishing_poly)))
        return out

    def free(self):
        if getattr(self, "handle", None) and getattr(self.ctx, "handle", None):
            self.ctx.lib.trp_domain_free(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
