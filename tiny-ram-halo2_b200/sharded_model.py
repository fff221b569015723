"""create_proof hot path sharded over the GPUs of one box (BASELINE.json configs[4]): one process per GPU,
torch.distributed (NCCL) only where the path has a real exchange.

  commit phase   the per-proof columns are round-robin over ranks (parallel.shard_columns): commit_lagrange (MSM of n + 1
                 points) and lagrange_to_coeff run on the owner; the 96-byte commitments are all_gathered.
  exchange       ONE all_gather of the coefficient-form columns (n_cols * n * 32 B in total over NVLink).
  quotient       the j - 1 cosets that determine h(X) are round-robin over ranks (parallel.shard_cosets): coset NTT of every
                 column + the quotient program on the owner; the n-value results are all_gathered and every rank recovers the
                 coefficients of h (trp_dev_cosets_to_coeff).
  commit h       the j - 1 pieces of h and the random polynomial are round-robin over ranks again.

Columns are synthetic (TinyRAM shape, tinyram_shape.py) and generated per column from (seed, column index), so that any
rank -- or a single GPU holding everything -- produces the same data and the sharded result can be compared bit for bit
with the one-GPU result.  Lookup / permutation products and the opening are per-column or sequential work that the
single-GPU model (prover_model.py) covers; they are not repeated here."""
from __future__ import annotations

import ctypes

import numpy as np

from . import parallel as PL
from . import poly as P
from . import synthetic, tinyram_shape
from ._lib import Q_CONTIGUOUS


class ShardedProverModel:
    def __init__(self, ctx, k: int, stream, dist=None, seed: int = 40, scale: float = 1.0, msm_batch: int = 32, streamed: bool = None,
                 block_slots: int = 8):
        import torch
        from .domain import EvaluationDomain
        self.torch, self.ctx, self.k, self.n, self.stream, self.dist = torch, ctx, k, 1 << k, stream, dist
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        self.msm_batch = msm_batch
        # streamed: the coefficient columns are all_gathered in blocks and transformed straight into the owner's coset buffer,
        # so no rank ever holds all n_cols coefficient columns (what lets k = 22 fit: 62 GiB of coefficients there); costs one
        # pass over the exchange per coset a rank owns.  Default: on from k = 21.
        self.streamed = (k >= 21) if streamed is None else streamed
        self.block_slots = block_slots
        self.dom = EvaluationDomain(ctx, 6, k)
        self.cosets = self.dom.j - 1
        self.shape = tinyram_shape.build(seed, scale=scale)
        self.ev = P.new_evaluator(ctx)
        self.prog = P.compile_ast(self.shape.ast, self.ev.modulus)
        g = self.shape.groups
        per_proof = ["advice", "instance", "permutation_z", "lookup_permuted_input", "lookup_permuted_table", "lookup_z"]
        self.proof_cols = [c for name in per_proof for c in range(g[name][0], g[name][0] + g[name][1])]
        proof_set = set(self.proof_cols)
        self.keygen_cols = [c for c in range(self.shape.n_columns) if c not in proof_set]
        self.n_proof = len(self.proof_cols)
        self.mine = PL.shard_columns(self.n_proof, self.world, self.rank)
        n, dev = self.n, "cuda"
        lib = ctx.lib
        pts = torch.empty((n + 1, 8), dtype=torch.int64, device=dev)
        torch.cuda.synchronize()
        synthetic.device_points(ctx, n + 1, pts.data_ptr())
        self.h_lagrange, self.h_g = ctypes.c_void_p(), ctypes.c_void_p()
        ctx.check(lib.trp_dev_bases_load(ctx.handle, pts.data_ptr(), n + 1, ctypes.byref(self.h_lagrange)))
        ctx.check(lib.trp_dev_bases_load(ctx.handle, pts.data_ptr(), n, ctypes.byref(self.h_g)))
        ctx.sync()
        del pts
        gen = torch.Generator(device=dev)

        def column(tag, c, out):
            gen.manual_seed((seed * 1000003 + tag * 100003 + c) & 0x7FFFFFFFFFFFFFFF)
            out.random_(0, 1 << 62, generator=gen)          # below 2^254: a valid Montgomery representation

        # this rank's per-proof columns (Lagrange basis) ++ blind row, staged for the MSM
        self.lag = torch.empty((max(len(self.mine), 1), n + 1, 4), dtype=torch.int64, device=dev)
        for j, c in enumerate(self.mine):
            column(1, c, self.lag[j])
        # keygen-time columns are static: every rank keeps them, already evaluated on each coset it owns
        self.my_cosets = PL.shard_cosets(self.cosets, self.world, self.rank)
        self.keygen_coset = torch.empty((len(self.keygen_cols), n, 4), dtype=torch.int64, device=dev)
        for j, c in enumerate(self.keygen_cols):
            column(2, c, self.keygen_coset[j])
        self.rand_poly = torch.empty((n, 4), dtype=torch.int64, device=dev)
        column(3, 0, self.rand_poly)
        self.coeff_local = torch.empty((max(len(self.mine), 1), n, 4), dtype=torch.int64, device=dev)
        self.coset_buf = torch.empty((self.n_proof, n, 4), dtype=torch.int64, device=dev)
        self.commit_local = torch.zeros((max(len(self.mine), 1), 12), dtype=torch.int64, device=dev)
        ptrs = [0] * self.shape.n_columns
        for j, c in enumerate(self.proof_cols):
            ptrs[c] = self.coset_buf[j].data_ptr()
        for j, c in enumerate(self.keygen_cols):
            ptrs[c] = self.keygen_coset[j].data_ptr()
        self.col_ptrs = ptrs
        torch.cuda.synchronize()

    # ---- the two callables of parallel.quotient_cosets_sharded ------------------------------------------------------------
    def _eval_coset(self, all_coeff, cs):
        ctx, lib, torch = self.ctx, self.ctx.lib, self.torch
        out = torch.empty((self.n, 4), dtype=torch.int64, device="cuda")
        ctx.check(lib.trp_dev_coeff_to_coset(self.dom.handle, all_coeff.data_ptr(), self.coset_buf.data_ptr(), self.n_proof, cs))
        self.ev.evaluate_device(self.prog, self.dom, self.col_ptrs, out.data_ptr(), coset=cs | Q_CONTIGUOUS)
        return out

    def _quotient_streamed(self):
        """per round r every rank evaluates coset r * world + rank (if it exists): the columns arrive as blocks of
        `block_slots` slots from every rank (one all_gather per block, all ranks take part in every collective) and each block
        is transformed into the coset buffer at once, so a rank holds its own coefficients, one block and the coset buffer"""
        torch, ctx, lib, dist = self.torch, self.ctx, self.ctx.lib, self.dist
        world, rank, n = self.world, self.rank, self.n
        m = len(self.mine)
        out = []
        for rnd in range((self.cosets + world - 1) // world):
            cs = rnd * world + rank
            active = cs < self.cosets
            for g0, cols in PL.all_gather_column_blocks(self.coeff_local[:m], self.n_proof, self.block_slots, dist):
                if active:
                    ctx.check(lib.trp_dev_coeff_to_coset(self.dom.handle, cols.data_ptr(), self.coset_buf[g0].data_ptr(), cols.shape[0], cs))
                    ctx.sync()
            if active:
                res = torch.empty((n, 4), dtype=torch.int64, device="cuda")
                self.ev.evaluate_device(self.prog, self.dom, self.col_ptrs, res.data_ptr(), coset=cs | Q_CONTIGUOUS)
                ctx.sync()
                out.append(res)
        return out

    def _combine(self, vals):
        ctx, lib, torch = self.ctx, self.ctx.lib, self.torch
        vals = vals.contiguous().clone()                     # consumed by the call
        h = torch.empty((self.cosets, self.n, 4), dtype=torch.int64, device="cuda")
        ctx.check(lib.trp_dev_cosets_to_coeff(self.dom.handle, vals.data_ptr(), self.cosets, h.data_ptr(), 1))
        self._keep = vals
        return h

    def prove_once(self):
        """One pass; returns ({phase: device ms on this rank}, commitments (n_proof, 12), h coefficients (j - 1, n, 4),
        h-piece commitments (j, 12)) -- the last three identical on every rank."""
        torch, ctx, lib, n, st, dist = self.torch, self.ctx, self.ctx.lib, self.n, self.stream, self.dist
        marks = []

        def mark(name):
            e = torch.cuda.Event(enable_timing=True)
            e.record(st)
            marks.append((name, e))

        m = len(self.mine)
        with torch.cuda.stream(st):
            mark("start")
            for b0 in range(0, m, self.msm_batch):
                nb = min(self.msm_batch, m - b0)
                ctx.check(lib.trp_dev_msm_batch(ctx.handle, self.h_lagrange, self.lag[b0].data_ptr(), n + 1, nb,
                                                self.commit_local[b0].data_ptr()))
            mark("commit_lagrange")
            if m:
                self.coeff_local[:m].copy_(self.lag[:m, :n])
                ctx.check(lib.trp_dev_lagrange_to_coeff(self.dom.handle, self.coeff_local.data_ptr(), m))
            mark("lagrange_to_coeff")
            commitments = PL.all_gather_columns(self.commit_local[:m], self.n_proof, dist)
            if self.streamed and self.world > 1:
                local = self._quotient_streamed()
                mark("quotient_cosets_streamed")
            else:
                all_coeff = PL.all_gather_columns(self.coeff_local[:m], self.n_proof, dist).contiguous()
                mark("all_gather_columns")
                local = [self._eval_coset(all_coeff, cs) for cs in self.my_cosets]
                del all_coeff
                mark("quotient_cosets")
            stacked = torch.stack(local) if local else torch.zeros((0, n, 4), dtype=torch.int64, device="cuda")
            h = self._combine(PL.all_gather_columns(stacked, self.cosets, dist))
            mark("gather_and_cosets_to_coeff")
            # commit the j - 1 pieces and the random polynomial, round-robin over ranks
            pieces = torch.cat([h, self.rand_poly[None]])
            mine_h = PL.shard_columns(self.cosets + 1, self.world, self.rank)
            hc_local = torch.zeros((max(len(mine_h), 1), 12), dtype=torch.int64, device="cuda")
            if mine_h:
                sel = pieces[mine_h].contiguous()
                ctx.check(lib.trp_dev_msm_batch(ctx.handle, self.h_g, sel.data_ptr(), n, len(mine_h), hc_local.data_ptr()))
            h_commit = PL.all_gather_columns(hc_local[:len(mine_h)], self.cosets + 1, dist)
            mark("commit_h")
        torch.cuda.synchronize()
        out = {}
        for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
            out[name + "_ms"] = a.elapsed_time(b)
        out["total_ms"] = marks[0][1].elapsed_time(marks[-1][1])
        return out, commitments, h, h_commit

    def close(self):
        lib = self.ctx.lib
        lib.trp_bases_free(self.h_lagrange); lib.trp_bases_free(self.h_g)
        self.dom.free()
