"""Mirror of halo2_proofs::plonk::permutation::prover::Argument::commit (plonk/permutation/prover.rs, halo2_proofs 0.2.0):
the grand-product columns Z of the permutation argument, one per chunk of cs_degree - 2 columns, computed on the GPU
(trp_permutation_product: term products, batch inversion and the prefix product in HBM).  The reference enables equality
on 188 columns (/root/reference/src/circuits/tables/prog.rs:151-152) => 47 Z columns at degree 6."""
from __future__ import annotations

import ctypes

import numpy as np

from ._lib import as_u64, ptr

# pasta_curves FieldExt::DELTA = 5^(2^32), Montgomery limbs are produced by the caller's field mul; we only need canonical ints
_MODULUS = {1: 0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001,    # ctx.curve VESTA  -> scalar field Fp
            0: 0x40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001}    # ctx.curve PALLAS -> scalar field Fq


def _limbs(v):
    return np.array([(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)], dtype=np.uint64)


def _to_int(l):
    return int(l[0]) | int(l[1]) << 64 | int(l[2]) << 128 | int(l[3]) << 192


def commit(domain, values, permutations, beta, gamma, chunk_len, blinding_factors, rand, after_chunk=None):
    """values / permutations: (m, n, 4) uint64 Montgomery columns (the column's Lagrange values / its sigma polynomial);
    beta, gamma: canonical Python ints; rand() returns a canonical int (the caller's RNG, drawn for the last
    `blinding_factors` rows of every Z in order); after_chunk(z), if given, runs after each Z is complete.  Returns the list
    of Z columns, each (n, 4) Montgomery."""
    ctx = domain.ctx
    p = _MODULUS[ctx.curve]
    R = (1 << 256) % p
    mont = lambda v: _limbs(v * R % p)
    values, permutations = as_u64(values), as_u64(permutations)
    m, n = values.shape[0], domain.n
    if values.shape != permutations.shape or values.shape[1] != n:
        raise ValueError("values and permutations must both be (columns, n, 4)")
    delta = pow(5, 1 << 32, p)
    deltaomega, last_z, sets = 1, None, []
    for lo in range(0, m, chunk_len):
        cols = list(range(lo, min(lo + chunk_len, m)))
        dbeta = np.stack([mont(deltaomega * pow(delta, c - lo, p) % p * beta % p) for c in cols])
        deltaomega = deltaomega * pow(delta, len(cols), p) % p
        vptr = (ctypes.c_void_p * len(cols))(*[values[c].ctypes.data for c in cols])
        sptr = (ctypes.c_void_p * len(cols))(*[permutations[c].ctypes.data for c in cols])
        z = np.empty((n, 4), dtype=np.uint64)
        ctx.check(ctx.lib.trp_permutation_product(domain.handle, vptr, sptr, len(cols), ptr(mont(beta)), ptr(mont(gamma)), ptr(dbeta),
                                                  ptr(last_z), ptr(z)))
        for i in range(n - blinding_factors, n):
            z[i] = mont(rand())
        last_z = z[n - (blinding_factors + 1)].copy()
        sets.append(z)
        if after_chunk is not None:          # create_proof draws the chunk's blind and commits before the next chunk starts
            after_chunk(z)
    return sets
