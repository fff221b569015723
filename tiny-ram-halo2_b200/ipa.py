"""Mirror of halo2_proofs::poly::commitment::prover::create_proof (poly/commitment/prover.rs, halo2_proofs 0.2.0): the
inner-product-argument opening that ends every create_proof of the reference (/root/reference/src/test_utils.rs:41,96).
Every vector stays on the GPU; per round the host sees two points (L_j, R_j) and one challenge, exactly what the
transcript needs.  Also mirrors arithmetic::{eval_polynomial, compute_inner_product, kate_division}.

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
    """The part of poly::commitment::Params the opening needs: g (n points), w, u; g ++ [w] is also loaded as MSM bases with
    the precomputed window table (Params::commit of the blinding polynomial S)."""

    def __init__(self, ctx, k, g, w, u):
        import torch
        self.ctx, self.k, self.n = ctx, k, 1 << k
        g = as_u64(g).reshape(-1, 8)
        if len(g) != self.n:
            raise ValueError("g must hold 2^k points")
        self.w, self.u = as_u64(w).reshape(8), as_u64(u).reshape(8)
        self.d_g = torch.from_numpy(g.view(np.int64)).cuda()
        self.d_uw = torch.from_numpy(np.stack([self.u, self.w]).view(np.int64)).cuda()
        h = ctypes.c_void_p()
        gw = torch.from_numpy(np.concatenate([g, self.w.reshape(1, 8)]).view(np.int64)).cuda()
        torch.cuda.synchronize()
        ctx.check(ctx.lib.trp_dev_bases_load(ctx.handle, gw.data_ptr(), self.n + 1, ctypes.byref(h)))
        ctx.sync()
        self.h_gw = h

    def free(self):
        if getattr(self, "h_gw", None) and getattr(self.ctx, "handle", None):
            self.ctx.lib.trp_bases_free(self.h_gw)
        self.h_gw = None


def create_proof(params: IpaParams, rand, transcript, p_poly, p_blind, x_3, rand_vector=None):
    """poly::commitment::prover::create_proof(params, rng, transcript, p_poly, p_blind, x_3).

    p_poly: (n, 4) Montgomery host array or a cuda int64 tensor (coefficient form); p_blind, x_3: canonical ints.
    rand() draws one canonical scalar; rand_vector(n), if given, draws n at once as an (n, 4) Montgomery array (the
    coefficients of the blinding polynomial S).  transcript: write_point((8,) affine Montgomery limbs),
    write_scalar(int), squeeze_challenge_scalar() -> int.  Nothing is returned: like halo2, the proof is what was written
    to the transcript."""
    import torch
    ctx, lib, n, k = params.ctx, params.ctx.lib, params.n, params.k
    p = _MODULUS[ctx.curve]
    R = (1 << 256) % p
    Rinv = pow(R, -1, p)
    mont = lambda v: _limbs(v % p * R % p)
    unmont = lambda l: _to_int(l) * Rinv % p
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()

    def evaluate(d_poly, x):
        out = torch.zeros(4, dtype=torch.int64, device="cuda")
        ctx.check(lib.trp_dev_eval_polynomials(ctx.handle, 0, d_poly.data_ptr(), n, n, 1, ptr(mont(x)), out.data_ptr()))
        ctx.sync()
        return unmont(out.cpu().numpy().view(np.uint64))

    def set_elem(d_vec, i, v):
        d_vec[i] = dev(mont(v))

    d_p = p_poly if hasattr(p_poly, "data_ptr") else dev(as_u64(p_poly))
    if d_p.numel() != 4 * n:
        raise ValueError("p_poly.len() != params.n")
    torch.cuda.synchronize()
    # random polynomial S with a root at x_3
    s_host = rand_vector(n) if rand_vector else np.stack([mont(rand()) for _ in range(n)])
    d_s = dev(as_u64(s_host))
    torch.cuda.synchronize()
    s0 = unmont(as_u64(s_host)[0])
    s_at_x3 = evaluate(d_s, x_3)
    set_elem(d_s, 0, s0 - s_at_x3)
    s_poly_blind = rand()
    # params.commit(&s_poly, s_poly_blind)
    d_sc = torch.cat([d_s.reshape(n, 4), dev(mont(s_poly_blind)).reshape(1, 4)])
    d_pt = torch.zeros((4, 12), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctx.check(lib.trp_dev_msm_batch(ctx.handle, params.h_gw, d_sc.data_ptr(), n + 1, 1, d_pt.data_ptr()))
    ctx.sync()
    transcript.write_point(d_pt[0].cpu().numpy().view(np.uint64)[:8].copy())
    xi = transcript.squeeze_challenge_scalar()
    z = transcript.squeeze_challenge_scalar()
    # P' = P - [v] G_0 + [xi] S
    d_pp = torch.empty_like(d_s)
    d_xi = dev(mont(xi))          # must outlive the kernel that reads it: the library's stream is invisible to torch's allocator
    torch.cuda.synchronize()
    ctx.check(lib.trp_dev_field_op(ctx.handle, 0, 2 | 16, d_s.data_ptr(), d_xi.data_ptr(), d_pp.data_ptr(), n))
    ctx.check(lib.trp_dev_field_op(ctx.handle, 0, 0, d_pp.data_ptr(), d_p.data_ptr(), d_pp.data_ptr(), n))
    ctx.sync()
    v = evaluate(d_pp, x_3)
    ctx.sync()
    pp0 = unmont(d_pp.reshape(n, 4)[0].cpu().numpy().view(np.uint64))
    set_elem(d_pp.reshape(n, 4), 0, pp0 - v)
    f = (s_poly_blind * xi + p_blind) % p
    d_b = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    # G' ++ [U, W]: L_j and R_j are computed as ONE batch of two MSMs over (G'_lo | G'_hi | U | W) with the scalar columns
    # (p'_hi | 0 | z <p'_hi, b_lo> | l_rand) and (0 | p'_lo | z <p'_lo, b_hi> | r_rand): one launch sequence per round
    # instead of four MSMs and two point additions (zero scalars cost nothing: zero digits are never sorted into buckets)
    d_g = torch.cat([params.d_g, params.d_uw])
    d_sc = torch.zeros((2, n + 2, 4), dtype=torch.int64, device="cuda")
    d_ip = torch.zeros((2, 4), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctx.check(lib.trp_dev_powers(ctx.handle, 0, ptr(mont(x_3)), n, d_b.data_ptr()))
    d_pp = d_pp.reshape(n, 4)
    for j in range(k):
        half = 1 << (k - j - 1)
        cur = 2 * half
        el = 32 * half                      # bytes per half vector of scalars
        ctx.check(lib.trp_dev_inner_products(ctx.handle, 0, d_pp.data_ptr() + el, 0, d_b.data_ptr(), 0, half, 1, d_ip[0].data_ptr()))
        ctx.check(lib.trp_dev_inner_products(ctx.handle, 0, d_pp.data_ptr(), 0, d_b.data_ptr() + el, 0, half, 1, d_ip[1].data_ptr()))
        ctx.sync()
        value_l, value_r = (unmont(r) for r in d_ip.cpu().numpy().view(np.uint64))
        l_rand, r_rand = rand(), rand()
        cols = d_sc.reshape(-1)[:2 * (cur + 2) * 4].reshape(2, cur + 2, 4)
        cols.zero_()
        cols[0, :half] = d_pp[half:cur]
        cols[1, half:cur] = d_pp[:half]
        cols[:, cur:] = dev(np.stack([mont(value_l * z), mont(l_rand), mont(value_r * z), mont(r_rand)])).reshape(2, 2, 4)
        torch.cuda.synchronize()
        ctx.check(lib.trp_dev_msm_var(ctx.handle, d_g.data_ptr(), cols.data_ptr(), cur + 2, 2, d_pt.data_ptr()))
        ctx.sync()
        pts = [r[:8].copy() for r in d_pt[:2].cpu().numpy().view(np.uint64)]
        transcript.write_point(pts[0])
        transcript.write_point(pts[1])
        u_j = transcript.squeeze_challenge_scalar()
        u_j_inv = pow(u_j, -1, p)
        ctx.check(lib.trp_dev_fold(ctx.handle, 0, d_pp.data_ptr(), half, ptr(mont(u_j_inv))))
        ctx.check(lib.trp_dev_fold(ctx.handle, 0, d_b.data_ptr(), half, ptr(mont(u_j))))
        ctx.check(lib.trp_dev_generator_collapse(ctx.handle, d_g.data_ptr(), half, ptr(mont(u_j))))
        ctx.sync()
        d_g[half:half + 2] = params.d_uw    # U, W follow the collapsed generators
        f = (f + l_rand * u_j_inv + r_rand * u_j) % p
    ctx.sync()
    c = unmont(d_pp[0].cpu().numpy().view(np.uint64))
    transcript.write_scalar(c)
    transcript.write_scalar(f)
