"""Mirror of halo2_proofs::plonk::lookup::prover (plonk/lookup/prover.rs, halo2_proofs 0.2.0): permute_expression_pair and
Permuted::commit_product on the GPU.  The reference's circuit has 31 lookups (/root/reference/src/circuits/tables/
even_bits.rs:158-170, out_table.rs:33-74, shift.rs:142-165, circuits/mod.rs:52-57)."""
from __future__ import annotations

import ctypes

import numpy as np

from ._lib import as_u64, ptr
from .permutation import _MODULUS, _limbs


class ConstraintSystemFailure(ValueError):
    """halo2's Error::ConstraintSystemFailure: an input value of the lookup does not occur in the table."""


def permute_expression_pair(ctx, input_expression, table_expression, usable_rows):
    """Returns (permuted_input, permuted_table) over the usable rows ((usable_rows, 4) Montgomery each); the caller
    appends its blinding rows, as create_proof does with its RNG."""
    a, s = as_u64(input_expression), as_u64(table_expression)
    if a.shape[0] < usable_rows or s.shape[0] < usable_rows:
        raise ValueError("expressions are shorter than usable_rows")
    pa = np.empty((usable_rows, 4), dtype=np.uint64)
    ps = np.empty((usable_rows, 4), dtype=np.uint64)
    ok = ctypes.c_int(1)
    ctx.check(ctx.lib.trp_permute_expression_pair(ctx.handle, ptr(a), ptr(s), usable_rows, ptr(pa), ptr(ps), ctypes.byref(ok)))
    if not ok.value:
        raise ConstraintSystemFailure("lookup input value not present in the table")
    return pa, ps


def commit_product(domain, compressed_input, compressed_table, permuted_input, permuted_table, beta, gamma, blinding_factors, rand):
    """Z of one lookup: n - blinding_factors computed rows followed by blinding_factors values from rand()."""
    ctx = domain.ctx
    p = _MODULUS[ctx.curve]
    R = (1 << 256) % p
    mont = lambda v: _limbs(v * R % p)
    n = domain.n
    cols = [as_u64(c) for c in (compressed_input, compressed_table, permuted_input, permuted_table)]
    if any(c.shape != (n, 4) for c in cols):
        raise ValueError("all four columns must hold n field elements")
    z = np.empty((n, 4), dtype=np.uint64)
    ctx.check(ctx.lib.trp_lookup_product(domain.handle, ptr(cols[0]), ptr(cols[1]), ptr(cols[2]), ptr(cols[3]), ptr(mont(beta)),
                                         ptr(mont(gamma)), ptr(z), n - blinding_factors))
    for i in range(n - blinding_factors, n):
        z[i] = mont(rand())
    return z
