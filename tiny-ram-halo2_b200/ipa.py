"""Mirror of halo2_proofs::poly::commitment::prover::create_proof (poly/commitment/prover.rs, halo2_proofs 0.2.0): the
inner-product-argument opening that ends every create_proof of the reference (/root/reference/src/test_utils.rs:41,96).
Every vector stays on the GPU; per round the host sees two points (L_j, R_j) and one challenge, exactly what the
transcript needs.  The rounds never collapse the generators: L_j and R_j are fixed-base MSMs over the ORIGINAL generators'
resident window table with the scalars p'[.] * s_t (trp_dev_ipa_round_scalars), which removes halo2's
parallel_generator_collapse from the prover altogether (measured at k = 20 on one B200: 165 ms -> see profiles/ipa_r02.md).  Also mirrors arithmetic::{eval_polynomial, compute_inner_product, kate_division}.

Device memory is held in torch tensors (int64 views of the 4 x u64 limbs); the arithmetic is libtrp.so's."""
from __future__ import annotations

import ctypes

import numpy as np

from ._lib import as_u64, ptr
from .permutation import _MODULUS, _limbs, _to_int


def eval_polynomial(ctx, poly, point):
    """arithmetic::eval_polynomial(poly, point): poly (n, 4) Montgomery host array, point (4,) Montgomery -> (4,)"""
    poly = as_u64(poly); out = np.zeros(4, dtype=np.uint64)
    ctx.check(ctx.lib.trp_eval_polynomial(ctx.handle, 0, ptr(poly), poly.size // 4, ptr(as_u64(point)), ptr(out)))
    return out


def compute_inner_product(ctx, a, b):
    """arithmetic::compute_inner_product(a, b); halo2 asserts equal lengths"""
    a, b = as_u64(a), as_u64(b)
    if a.shape != b.shape:
        raise ValueError("compute_inner_product: a.len() != b.len()")
    out = np.zeros(4, dtype=np.uint64)
    ctx.check(ctx.lib.trp_compute_inner_product(ctx.handle, 0, ptr(a), ptr(b), a.size // 4, ptr(out)))
    return out


def kate_division(ctx, a, b):
    """arithmetic::kate_division(a, b): coefficients of (a(X) - a(b)) / (X - b), one fewer than a"""
    a = as_u64(a); n = a.size // 4
    q = np.zeros((max(n - 1, 0), 4), dtype=np.uint64)
    ctx.check(ctx.lib.trp_kate_division(ctx.handle, 0, ptr(a), n, ptr(as_u64(b)), ptr(q)))
    return q


class IpaParams:
    """The part of poly::commitment::Params the opening needs: g (n points), w, u.  g ++ [u, w] is loaded as MSM bases with the
    precomputed window table: Params::commit of the blinding polynomial S and every round's L_j / R_j run over it."""

    def __init__(self, ctx, k, g, w, u):
        import torch
        self.ctx, self.k, self.n = ctx, k, 1 << k
        ctx.bind_torch_stream()              # torch tensor operations and trp_dev_* calls are mixed below: one stream for both
        g = as_u64(g).reshape(-1, 8)
        if len(g) != self.n:
            raise ValueError("g must hold 2^k points")
        self.w, self.u = as_u64(w).reshape(8), as_u64(u).reshape(8)
        self.g_host = g
        self.d_g = torch.from_numpy(g.view(np.int64)).cuda()
        self.d_uw = torch.from_numpy(np.stack([self.u, self.w]).view(np.int64)).cuda()
        self._tables = {}
        self.h_guw = self.table(0, 1)

    def table(self, rank, world):
        """window table over g[lo:hi] ++ [u, w], [lo, hi) = rank's slice of the generators (the whole range for world = 1)"""
        import torch
        key = (rank, world)
        if key not in self._tables:
            lo, hi = rank * self.n // world, (rank + 1) * self.n // world
            pts = torch.cat([self.d_g.reshape(self.n, 8)[lo:hi], self.d_uw.reshape(2, 8)])
            h = ctypes.c_void_p()
            torch.cuda.current_stream().synchronize()
            self.ctx.check(self.ctx.lib.trp_dev_bases_load(self.ctx.handle, pts.data_ptr(), hi - lo + 2, ctypes.byref(h)))
            self.ctx.sync()
            self._tables[key] = h
        return self._tables[key]

    def free(self):
        if getattr(self.ctx, "handle", None):
            for h in getattr(self, "_tables", {}).values():
                self.ctx.lib.trp_bases_free(h)
        self._tables = {}
        self.h_guw = None


def create_proof(params: IpaParams, rand, transcript, p_poly, p_blind, x_3, rand_vector=None, dist=None, trace=None):
    """poly::commitment::prover::create_proof(params, rng, transcript, p_poly, p_blind, x_3).

    p_poly: (n, 4) Montgomery host array or a cuda int64 tensor (coefficient form); p_blind, x_3: canonical ints.
    rand() draws one canonical scalar; rand_vector(n), if given, draws n at once as an (n, 4) Montgomery array or cuda tensor
    (the coefficients of the blinding polynomial S).  transcript: write_point((8,) affine Montgomery limbs),
    write_scalar(int), squeeze_challenge_scalar() -> int.  Nothing is returned: like halo2, the proof is what was written
    to the transcript.

    Round j (cur = n / 2^j live entries of p' and b, s = the 2^j challenge products):
        L_j = sum_b colL[b] G_b + [z <p'_hi, b_lo>] U + [l_rand] W,   colL[t cur + i] = p'[half + i] s_t  (i < half), 0 otherwise
        R_j = sum_b colR[b] G_b + [z <p'_lo, b_hi>] U + [r_rand] W,   colR[t cur + half + i] = p'[i] s_t
    as ONE batch of two fixed-base MSMs over the table of g ++ [u, w]; then p' and b are folded and s doubles.  These are the
    group elements halo2 computes from its collapsed G' (G'_i = sum_t s_t G_{t cur + i}), so the transcript is identical.

    dist (an initialised process group, every rank calling with the same inputs): each large MSM's POINT RANGE is split --
    rank r owns the generators [r n / G, (r + 1) n / G) (its own window table) and the partial sums are all_gathered and added
    (trp_dev_points_sum), best_multiexp's own combination step; p', b and s (a few MiB) stay replicated."""
    import torch
    ctx, lib, n, k = params.ctx, params.ctx.lib, params.n, params.k
    ctx.bind_torch_stream()
    p = _MODULUS[ctx.curve]
    R = (1 << 256) % p
    Rinv = pow(R, -1, p)
    mont = lambda v: _limbs(v % p * R % p)
    unmont = lambda l: _to_int(l) * Rinv % p
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()
    import time as _time

    def mark(label):                        # development aid (tests/gpu_ipa_trace.py): drain the stream and take a timestamp
        if trace is not None:
            torch.cuda.synchronize()
            trace.append((label, _time.perf_counter()))

    mark("start")
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    if n % world:
        raise ValueError("the number of ranks must divide n")

    def evaluate(d_poly, x):
        out = torch.zeros(4, dtype=torch.int64, device="cuda")
        ctx.check(lib.trp_dev_eval_polynomials(ctx.handle, 0, d_poly.data_ptr(), n, n, 1, ptr(mont(x)), out.data_ptr()))
        return unmont(out.cpu().numpy().view(np.uint64))

    def set_elem(d_vec, i, v):
        d_vec[i] = dev(mont(v))

    d_p = p_poly if hasattr(p_poly, "data_ptr") else dev(as_u64(p_poly))
    if d_p.numel() != 4 * n:
        raise ValueError("p_poly.len() != params.n")
    # this rank's slice of the generators (all of them on one GPU) and its table
    lo, hi = rank * n // world, (rank + 1) * n // world
    cnt = hi - lo
    table = params.table(rank, world)
    cols = torch.zeros((2, cnt + 2, 4), dtype=torch.int64, device="cuda")       # scalars of g[lo:hi] ++ [u, w], one column per point
    d_pt = torch.zeros((2, 12), dtype=torch.int64, device="cuda")
    gathered = torch.zeros((world, 2, 12), dtype=torch.int64, device="cuda") if world > 1 else None

    def msm_pair(m):
        """the m columns of `cols` over the (sliced) table, partial sums combined: -> m affine points as (8,) limb arrays"""
        ctx.check(lib.trp_dev_msm_batch(ctx.handle, table, cols.data_ptr(), cnt + 2, m, d_pt.data_ptr()))
        if world > 1:
            dist.all_gather_into_tensor(gathered.reshape(-1), d_pt.reshape(-1))
            for c_ in range(m):
                part = gathered[:, c_].contiguous()
                ctx.check(lib.trp_dev_points_sum(ctx.handle, part.data_ptr(), world, d_pt[c_].data_ptr()))
        return [r_[:8].copy() for r_ in d_pt[:m].cpu().numpy().view(np.uint64)]

    # random polynomial S with a root at x_3
    s_host = rand_vector(n) if rand_vector else np.stack([mont(rand()) for _ in range(n)])
    mark("draw S")
    if hasattr(s_host, "data_ptr"):             # already on the device (a seed expanded there: trp_dev_random_field)
        d_s = s_host.reshape(n, 4)
        s0 = unmont(d_s[0].cpu().numpy().view(np.uint64))
    else:
        d_s = dev(as_u64(s_host))
        s0 = unmont(as_u64(s_host)[0])
    mark("upload S")
    s_at_x3 = evaluate(d_s, x_3)
    set_elem(d_s, 0, s0 - s_at_x3)
    s_poly_blind = rand()
    # params.commit(&s_poly, s_poly_blind): over this rank's slice, the blind's term on rank 0
    cols[0, :cnt] = d_s.reshape(n, 4)[lo:hi]
    if rank == 0:
        cols[0, cnt + 1] = dev(mont(s_poly_blind))
    transcript.write_point(msm_pair(1)[0])
    mark("commit S")
    xi = transcript.squeeze_challenge_scalar()
    z = transcript.squeeze_challenge_scalar()
    # P' = P - [v] G_0 + [xi] S
    d_pp = torch.empty_like(d_s)
    d_xi = dev(mont(xi))
    ctx.check(lib.trp_dev_field_op(ctx.handle, 0, 2 | 16, d_s.data_ptr(), d_xi.data_ptr(), d_pp.data_ptr(), n))
    ctx.check(lib.trp_dev_field_op(ctx.handle, 0, 0, d_pp.data_ptr(), d_p.data_ptr(), d_pp.data_ptr(), n))
    v = evaluate(d_pp, x_3)
    pp0 = unmont(d_pp.reshape(n, 4)[0].cpu().numpy().view(np.uint64))
    set_elem(d_pp.reshape(n, 4), 0, pp0 - v)
    f = (s_poly_blind * xi + p_blind) % p
    d_pp = d_pp.reshape(n, 4)
    d_b = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    ctx.check(lib.trp_dev_powers(ctx.handle, 0, ptr(mont(x_3)), n, d_b.data_ptr()))
    d_z = dev(mont(z))
    d_ip = torch.zeros((2, 4), dtype=torch.int64, device="cuda")
    s_cur = dev(mont(1)).reshape(1, 4)
    s_next = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    s_bufs = [torch.empty((n, 4), dtype=torch.int64, device="cuda"), s_next]
    mark("P' and v")
    cur = n
    for j in range(k):
        half = cur // 2
        el = 32 * half                      # bytes per half vector of scalars
        ctx.check(lib.trp_dev_inner_products(ctx.handle, 0, d_pp.data_ptr() + el, 0, d_b.data_ptr(), 0, half, 1, d_ip[0].data_ptr()))
        ctx.check(lib.trp_dev_inner_products(ctx.handle, 0, d_pp.data_ptr(), 0, d_b.data_ptr() + el, 0, half, 1, d_ip[1].data_ptr()))
        l_rand, r_rand = rand(), rand()
        ctx.check(lib.trp_dev_ipa_round_scalars(ctx.handle, 0, d_pp.data_ptr(), s_cur.data_ptr(), cur, lo, cnt, cnt + 2, cols.data_ptr()))
        if rank == 0:                       # [z <p', b>] U + [rand] W: one rank adds them (the values never visit the host)
            ctx.check(lib.trp_dev_field_op(ctx.handle, 0, 2 | 16, d_ip.data_ptr(), d_z.data_ptr(), d_ip.data_ptr(), 2))
            cols[:, cnt] = d_ip
            cols[:, cnt + 1] = dev(np.stack([mont(l_rand), mont(r_rand)]))
        else:
            cols[:, cnt:] = 0
        pts = msm_pair(2)
        mark(f"round {cur}")
        transcript.write_point(pts[0])
        transcript.write_point(pts[1])
        u_j = transcript.squeeze_challenge_scalar()
        u_j_inv = pow(u_j, -1, p)
        ctx.check(lib.trp_dev_fold(ctx.handle, 0, d_pp.data_ptr(), half, ptr(mont(u_j_inv))))
        ctx.check(lib.trp_dev_fold(ctx.handle, 0, d_b.data_ptr(), half, ptr(mont(u_j))))
        nxt = s_bufs[j & 1]
        ctx.check(lib.trp_dev_ipa_s_double(ctx.handle, 0, s_cur.data_ptr(), n // cur, ptr(mont(u_j)), nxt.data_ptr()))
        s_cur = nxt
        f = (f + l_rand * u_j_inv + r_rand * u_j) % p
        cur = half
    c = unmont(d_pp[0].cpu().numpy().view(np.uint64))
    transcript.write_scalar(c)
    transcript.write_scalar(f)
