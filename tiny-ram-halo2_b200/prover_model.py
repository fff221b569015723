"""create_proof@k workload model: replays, on device-resident synthetic columns
of the TinyRAM circuit's shape (Appendix B), every hot-path call halo2_proofs' create_proof makes for one proof --

  phase 2-5  497 x commit_lagrange (MSM of n + 1 points) and 497 x lagrange_to_coeff           (instance, advice,
             lookup A'/S', permutation Z, lookup Z columns); between them, on the device:
             theta-compression of the 31 lookups' input / table expressions (the quotient VM over the Lagrange basis),
             31 x permute_expression_pair, 47 x permutation grand product (chunks of 4 columns, chained), 31 x lookup
             grand product
  phase 7    for each of the 2^(extended_k - k) cosets: 497 x coeff_to_coset (one size-n NTT per column; the 214 fixed /
             sigma / selector columns are keygen-time data and enter already evaluated), then the quotient program over
             all 711 columns; extended_to_coeff with divide_by_vanishing_poly; commit of the 5 h pieces and of the random
             blinding polynomial (6 coefficient-basis MSMs of n points)

  phase 8-9   the evaluations at x and x * omega (eval_polynomial of every coefficient-form column),
             multiopen's three kate_divisions and one poly::commitment::prover::create_proof (k IPA rounds: two MSMs over
             the collapsing generators, two inner products, the folds and parallel_generator_collapse per round)

-- and reports device time per phase.  What it does NOT contain ("next" rows): witness synthesis, the
transcript hashing and multiopen's linear combinations of the opened polynomials.  The synthetic lookups use the compressed input, rotated by one row, as
their table (so that every input value occurs in the table, as in a satisfied circuit); the table expression is still
evaluated.  The reference itself only runs this
path at k <= 14 on CPU (src/test_utils.rs:20); k = 20 is BASELINE.json's target configuration.

Scalars follow the distribution note of SURVEY.md 8(a): instance / advice columns are 90 % {0,1}, 8 % < 2^32, 2 % full
width; permuted lookup columns and grand products are uniform."""
from __future__ import annotations

import ctypes
import time

import numpy as np

from . import poly as P
from . import synthetic, tinyram_shape
from ._lib import ptr, Q_CONTIGUOUS

_R2 = {1: [0x8c78ecb30000000f, 0xd7d30dbd8b0de0e7, 0x7797a99bc3c95d18, 0x096d41af7b9cb714],     # Fp (ctx curve VESTA)
       0: [0xfc9678ff0000000f, 0x67bb433d891a16e3, 0x7fae231004ccf590, 0x096d41af7ccfdaa9]}     # Fq (ctx curve PALLAS)


class _NullTranscript:
    """Challenges from a seeded stream; writes are dropped (the transcript hash stays on the host side of the boundary)."""

    def __init__(self, rng, p):
        self.rng, self.p = rng, p

    def write_point(self, P):
        pass

    def write_scalar(self, s):
        pass

    def squeeze_challenge_scalar(self):
        return self.rng.randrange(1, self.p)


class CreateProofModel:
    def __init__(self, ctx, k: int, stream, seed: int = 40, msm_batch: int = 32, scale: float = 1.0, reduced_cosets: bool = True):
        import torch
        self.torch, self.ctx, self.k, self.n, self.stream = torch, ctx, k, 1 << k, stream
        self.msm_batch = msm_batch
        lib = ctx.lib
        from .domain import EvaluationDomain
        self.dom = EvaluationDomain(ctx, 6, k)
        # the quotient has degree < 5n: 5 of the 8 cosets determine it (trp_dev_cosets_to_coeff); reduced_cosets=False
        # replays halo2's own data flow (all 2^(extended_k-k) cosets, then extended_to_coeff)
        self.reduced = reduced_cosets
        self.cosets = (self.dom.j - 1) if reduced_cosets else 1 << (self.dom.extended_k - k)
        self.shape = tinyram_shape.build(seed, scale=scale)
        self.ev = P.new_evaluator(ctx)
        self.prog = P.compile_ast(self.shape.ast, self.ev.modulus)
        g = self.shape.groups
        per_proof = ["advice", "instance", "permutation_z", "lookup_permuted_input", "lookup_permuted_table", "lookup_z"]
        self.proof_cols = [c for name in per_proof for c in range(g[name][0], g[name][0] + g[name][1])]
        self.keygen_cols = [c for c in range(self.shape.n_columns) if c not in set(self.proof_cols)]
        self.n_proof = len(self.proof_cols)
        n, dev = self.n, "cuda"
        gen = torch.Generator(device=dev); gen.manual_seed(seed)
        # ---- bases: g_lagrange ++ [w] (n + 1 points) and g (n points), synthetic progressions generated on the device ------
        pts = torch.empty((n + 1, 8), dtype=torch.int64, device=dev)
        torch.cuda.synchronize()
        synthetic.device_points(ctx, n + 1, pts.data_ptr())
        self.h_lagrange, self.h_g = ctypes.c_void_p(), ctypes.c_void_p()
        ctx.check(lib.trp_dev_bases_load(ctx.handle, pts.data_ptr(), n + 1, ctypes.byref(self.h_lagrange)))
        ctx.check(lib.trp_dev_bases_load(ctx.handle, pts.data_ptr(), n, ctypes.byref(self.h_g)))
        ctx.sync()
        del pts
        # ---- per-proof Lagrange columns --------------------------------------------------------------------------------------
        n_small = g["advice"][1] + g["instance"][1]          # TinyRAM-shaped columns come first in proof_cols
        self.lag = torch.empty((self.n_proof, n, 4), dtype=torch.int64, device=dev)
        r2 = torch.tensor(np.array(_R2[ctx.curve], dtype=np.uint64).view(np.int64), device=dev)
        for c in range(self.n_proof):
            col = self.lag[c]
            col.random_(0, 1 << 62, generator=gen)           # uniform, already a valid Montgomery representation
            if c < n_small:
                kind = torch.rand(n, device=dev, generator=gen)
                small = torch.zeros((n, 4), dtype=torch.int64, device=dev)
                small[:, 0] = torch.where(kind < 0.9, torch.randint(0, 2, (n,), device=dev, generator=gen),
                                          torch.randint(0, 1 << 32, (n,), device=dev, generator=gen))
                torch.cuda.synchronize()
                ctx.check(lib.trp_dev_field_op(ctx.handle, 0, 2 | 16, small.data_ptr(), r2.data_ptr(), small.data_ptr(), n))
                ctx.sync()
                col.copy_(torch.where((kind < 0.98)[:, None], small, col))
        self.blinds = torch.randint(0, 1 << 62, (self.n_proof, 4), dtype=torch.int64, device=dev, generator=gen)
        self.keygen_coset = torch.randint(0, 1 << 62, (len(self.keygen_cols), n, 4), dtype=torch.int64, device=dev, generator=gen)
        self.coset_buf = torch.empty((self.n_proof, n, 4), dtype=torch.int64, device=dev)
        self.stage = torch.empty((msm_batch, n + 1, 4), dtype=torch.int64, device=dev)
        self.commitments = torch.zeros((self.n_proof, 12), dtype=torch.int64, device=dev)
        self.h_ext = torch.empty((self.cosets, n, 4), dtype=torch.int64, device=dev)
        self.h_coeff = torch.randint(0, 1 << 62, (6, n, 4), dtype=torch.int64, device=dev, generator=gen)   # [5] = random poly
        self.h_commit = torch.zeros((6, 12), dtype=torch.int64, device=dev)
        self.coeff = torch.empty_like(self.lag)
        ptrs = [0] * self.shape.n_columns
        for j, c in enumerate(self.proof_cols):
            ptrs[c] = self.coset_buf[j].data_ptr()
        for j, c in enumerate(self.keygen_cols):
            ptrs[c] = self.keygen_coset[j].data_ptr()
        self.col_ptrs = ptrs
        # ---- row f1: lookup permutation and grand products -----------------------------------------------------------
        lag_ptrs = [0] * self.shape.n_columns           # Lagrange-basis view of every column (keygen columns: stand-ins)
        self.row_of = {c: j for j, c in enumerate(self.proof_cols)}
        for j, c in enumerate(self.proof_cols):
            lag_ptrs[c] = self.lag[j].data_ptr()
        for j, c in enumerate(self.keygen_cols):
            lag_ptrs[c] = self.keygen_coset[j].data_ptr()
        self.lag_ptrs = lag_ptrs
        self.lookup_progs = [(P.compile_ast(a, self.ev.modulus), P.compile_ast(t, self.ev.modulus)) for a, t in self.shape.lookup_exprs]
        self.n_lookups = len(self.lookup_progs)
        self.compressed = torch.empty((self.n_lookups, 2, n, 4), dtype=torch.int64, device=dev)
        self.blinding_factors = 5
        self.usable_rows = n - (self.blinding_factors + 1)
        p = self.ev.modulus
        R = (1 << 256) % p
        limbs = lambda v: np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
        import random as _random
        rr = _random.Random(seed)
        self.beta, self.gamma = rr.randrange(p), rr.randrange(p)
        self.beta_m, self.gamma_m = limbs(self.beta * R % p), limbs(self.gamma * R % p)
        delta = pow(5, 1 << 32, p)
        # ---- row f2: opening ---------------------------------------------------------------------------------------------------
        from . import ipa as _ipa
        self._ipa = _ipa
        self.open_rng = rr
        self.evals = torch.zeros((2 * self.n_proof + 8, 4), dtype=torch.int64, device=dev)
        self.q_poly = torch.empty((n, 4), dtype=torch.int64, device=dev)
        self.x_m = limbs(rr.randrange(p) * R % p)
        self.ipa_params = None
        self.perm_chunks = []
        pc = self.shape.perm_cols
        sig0 = g["permutation_sigma"][0]
        for ch, lo in enumerate(range(0, len(pc), tinyram_shape.PERM_CHUNK)):
            cols = pc[lo:lo + tinyram_shape.PERM_CHUNK]
            vptr = (ctypes.c_void_p * len(cols))(*[lag_ptrs[c] for c in cols])
            sptr = (ctypes.c_void_p * len(cols))(*[lag_ptrs[sig0 + lo + i] for i in range(len(cols))])
            dbeta = np.stack([limbs(pow(delta, lo + i, p) * self.beta % p * R % p) for i in range(len(cols))])
            self.perm_chunks.append((vptr, sptr, len(cols), dbeta))
        torch.cuda.synchronize()

    def describe(self):
        c = self.prog.counts()
        return {"k": self.k, "per_proof_columns": self.n_proof, "keygen_columns": len(self.keygen_cols),
                "expressions": self.shape.n_expressions, "program": c, "vm_registers": self.prog.n_regs, "cosets_evaluated": self.cosets,
                "cosets_in_extended_domain": 1 << (self.dom.extended_k - self.k)}

    def prove_once(self):
        """One pass over the hot path; returns {phase: device milliseconds} measured with CUDA events on the ctx stream."""
        torch, ctx, lib, n, st = self.torch, self.ctx, self.ctx.lib, self.n, self.stream
        ev = lambda: torch.cuda.Event(enable_timing=True)
        marks = []

        def mark(name):
            e = ev(); e.record(st); marks.append((name, e))

        g = self.shape.groups
        rows = lambda name: range(self.row_of[g[name][0]], self.row_of[g[name][0]] + g[name][1])

        def commit_rows(r):
            for b0 in range(r.start, r.stop, self.msm_batch):
                nb = min(self.msm_batch, r.stop - b0)
                self.stage[:nb, :n].copy_(self.lag[b0:b0 + nb])
                self.stage[:nb, n].copy_(self.blinds[b0:b0 + nb])
                ctx.check(lib.trp_dev_msm_batch(ctx.handle, self.h_lagrange, self.stage.data_ptr(), n + 1, nb,
                                                self.commitments[b0].data_ptr()))

        la, ls, lz, pz = rows("lookup_permuted_input"), rows("lookup_permuted_table"), rows("lookup_z"), rows("permutation_z")
        times = {}

        def timed(name, fn):
            a = ev(); a.record(st); fn(); b = ev(); b.record(st)
            times.setdefault(name, []).append((a, b))

        with torch.cuda.stream(st):
            mark("start")
            # phases 2-3: commitments of the instance / advice columns (MSM of n + 1 points each: column ++ blind)
            timed("commit_lagrange_ms", lambda: commit_rows(range(0, min(la.start, pz.start))))
            # phase 4 (theta): compress the lookups' expressions, permute them, commit A' and S'
            def compress():
                for i, (pi, pt) in enumerate(self.lookup_progs):
                    self.ev.evaluate_device(pi, self.dom, self.lag_ptrs, self.compressed[i, 0].data_ptr(), coset=0 | Q_CONTIGUOUS)
                    self.ev.evaluate_device(pt, self.dom, self.lag_ptrs, self.compressed[i, 1].data_ptr(), coset=0 | Q_CONTIGUOUS)
                    self.compressed[i, 1].copy_(torch.roll(self.compressed[i, 0], 1, 0))      # see module docstring
            timed("lookup_compress_ms", compress)
            def permute():
                ok = ctypes.c_int(1)
                u = self.usable_rows
                for i in range(self.n_lookups):
                    ctx.check(lib.trp_dev_permute_expression_pair(ctx.handle, self.compressed[i, 0].data_ptr(), self.compressed[i, 1].data_ptr(),
                                                                  u, self.lag[la.start + i].data_ptr(), self.lag[ls.start + i].data_ptr(),
                                                                  ctypes.byref(ok)))
                    self.lookups_ok = self.lookups_ok and bool(ok.value)
            self.lookups_ok = True
            timed("lookup_permute_ms", permute)
            timed("commit_lagrange_ms", lambda: (commit_rows(la), commit_rows(ls)))
            # phase 5 (beta, gamma): permutation and lookup grand products, commit the Z columns
            def perm_products():
                last = None
                for ch, (vptr, sptr, m, dbeta) in enumerate(self.perm_chunks):
                    z = self.lag[pz.start + ch]
                    ctx.check(lib.trp_dev_permutation_product(self.dom.handle, vptr, sptr, m, ptr(self.beta_m), ptr(self.gamma_m), ptr(dbeta),
                                                              last, z.data_ptr()))
                    last = z[self.usable_rows].data_ptr()
            timed("permutation_product_ms", perm_products)
            def lookup_products():
                for i in range(self.n_lookups):
                    ctx.check(lib.trp_dev_lookup_product(self.dom.handle, self.compressed[i, 0].data_ptr(), self.compressed[i, 1].data_ptr(),
                                                         self.lag[la.start + i].data_ptr(), self.lag[ls.start + i].data_ptr(),
                                                         ptr(self.beta_m), ptr(self.gamma_m), self.lag[lz.start + i].data_ptr(),
                                                         n - self.blinding_factors))
            timed("lookup_product_ms", lookup_products)
            timed("commit_lagrange_ms", lambda: (commit_rows(pz), commit_rows(lz)))
            mark("commit_and_products")
            self.coeff.copy_(self.lag)
            ctx.check(lib.trp_dev_lagrange_to_coeff(self.dom.handle, self.coeff.data_ptr(), self.n_proof))
            mark("lagrange_to_coeff")
            t_ntt = t_vm = 0.0
            spans = []
            for cs in range(self.cosets):
                a = ev(); a.record(st)
                ctx.check(lib.trp_dev_coeff_to_coset(self.dom.handle, self.coeff.data_ptr(), self.coset_buf.data_ptr(), self.n_proof, cs))
                b = ev(); b.record(st)
                if self.reduced:
                    self.ev.evaluate_device(self.prog, self.dom, self.col_ptrs, self.h_ext[cs].data_ptr(), coset=cs | Q_CONTIGUOUS)
                else:
                    self.ev.evaluate_device(self.prog, self.dom, self.col_ptrs, self.h_ext.data_ptr(), coset=cs)
                c = ev(); c.record(st)
                spans.append((a, b, c))
            mark("quotient_cosets")
            if self.reduced:
                ctx.check(lib.trp_dev_cosets_to_coeff(self.dom.handle, self.h_ext.data_ptr(), self.cosets, self.h_coeff.data_ptr(), 1))
            else:
                ctx.check(lib.trp_dev_extended_to_coeff(self.dom.handle, self.h_ext.data_ptr(), self.h_coeff.data_ptr(), 1))
            mark("extended_to_coeff")
            ctx.check(lib.trp_dev_msm_batch(ctx.handle, self.h_g, self.h_coeff.data_ptr(), n, 6, self.h_commit.data_ptr()))
            mark("commit_h")
            # phase 8: evaluations at x and x * omega (every per-proof column, coefficient form); multiopen quotients
            ctx.check(lib.trp_dev_eval_polynomials(ctx.handle, 0, self.coeff.data_ptr(), n, n, self.n_proof, ptr(self.x_m), self.evals.data_ptr()))
            ctx.check(lib.trp_dev_eval_polynomials(ctx.handle, 0, self.coeff.data_ptr(), n, n, self.n_proof, ptr(self.beta_m),
                                                   self.evals[self.n_proof].data_ptr()))
            ctx.check(lib.trp_dev_eval_polynomials(ctx.handle, 0, self.h_coeff.data_ptr(), n, n, 6, ptr(self.x_m), self.evals[2 * self.n_proof].data_ptr()))
            mark("evaluations")
            for i in range(3):
                ctx.check(lib.trp_dev_kate_division(ctx.handle, 0, self.coeff[i].data_ptr(), n, ptr(self.x_m), self.q_poly.data_ptr()))
            mark("kate_division")
            # phase 9: the IPA opening of the final polynomial
            if self.ipa_params is None:
                self.ipa_params = self._make_ipa_params()
            self._ipa.create_proof(self.ipa_params, lambda: self.open_rng.randrange(self.ev.modulus), _NullTranscript(self.open_rng, self.ev.modulus),
                                   self.coeff[0], 12345, 67890, rand_vector=lambda m: self.h_coeff[5].cpu().numpy().view(np.uint64))
            mark("ipa_open")
        torch.cuda.synchronize()
        out = {}
        for (_, e_prev), (name, e) in zip(marks[:-1], marks[1:]):
            out[name + "_ms"] = e_prev.elapsed_time(e)
        for name, pairs in times.items():
            out[name] = sum(a.elapsed_time(b) for a, b in pairs)
        out["coset_ntt_ms"] = sum(a.elapsed_time(b) for a, b, _ in spans)
        out["quotient_vm_ms"] = sum(b.elapsed_time(c) for _, b, c in spans)
        out["total_ms"] = marks[0][1].elapsed_time(marks[-1][1])
        return out

    def _make_ipa_params(self):
        """IpaParams over the model's synthetic generators (g = the first n points of the Lagrange bases' progression)."""
        torch, ctx, n = self.torch, self.ctx, self.n
        pts = torch.empty((n + 2, 8), dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        synthetic.device_points(ctx, n + 2, pts.data_ptr())
        ctx.sync()
        host = pts.cpu().numpy().view(np.uint64)
        return self._ipa.IpaParams(ctx, self.k, host[:n], host[n], host[n + 1])

    def close(self):
        lib = self.ctx.lib
        lib.trp_bases_free(self.h_lagrange); lib.trp_bases_free(self.h_g)
        self.dom.free()
