"""CPU checks of the oracle's restatement of SURVEY.md 8(f) row f1 (ff::BatchInvert, permutation::Argument::commit,
lookup::permute_expression_pair / commit_product): the defining properties the halo2 verifier relies on."""
import random

import pytest

from util import pm

FIELDS = [pm.Fp, pm.Fq]


@pytest.mark.parametrize("F", FIELDS, ids=["Fp", "Fq"])
def test_batch_invert_skips_zeros(F):
    rng = random.Random(1)
    vals = [rng.randrange(F.p) for _ in range(50)]
    vals[3] = vals[17] = 0
    out = pm.batch_invert(F, vals)
    for v, o in zip(vals, out):
        assert (o == 0) if v == 0 else (v * o % F.p == 1)
    assert pm.batch_invert(F, []) == [] and pm.batch_invert(F, [0, 0]) == [0, 0]


@pytest.mark.parametrize("F", FIELDS, ids=["Fp", "Fq"])
def test_permutation_product_closes_for_a_valid_permutation(F):
    """When the column values respect the permutation sigma, the grand product returns to 1 on the last usable row."""
    rng = random.Random(2)
    k, m, chunk, bf = 5, 5, 2, 3
    n = 1 << k
    omega = F.root_of_unity(k)
    usable = n - (bf + 1)
    # identity labels delta^c * omega^i; a random permutation of the (column, row) cells of the usable rows
    cells = [(c, i) for c in range(m) for i in range(usable)]
    shuffled = cells[:]
    rng.shuffle(shuffled)
    label = lambda c, i: F.pow(F.DELTA, c) * F.pow(omega, i) % F.p
    sigma = [[label(c, i) for i in range(n)] for c in range(m)]
    values = [[rng.randrange(F.p) for _ in range(n)] for _ in range(m)]
    # make cycles: cell -> shuffled cell; values constant on cycles
    perm = dict(zip(cells, shuffled))
    seen = set()
    for start in cells:
        if start in seen:
            continue
        v, cur = rng.randrange(F.p), start
        while cur not in seen:
            seen.add(cur)
            values[cur[0]][cur[1]] = v
            cur = perm[cur]
    for (c, i), (c2, i2) in perm.items():
        sigma[c][i] = label(c2, i2)
    beta, gamma = rng.randrange(F.p), rng.randrange(F.p)
    zs = pm.permutation_commit(F, omega, n, values, sigma, beta, gamma, chunk, bf, lambda: rng.randrange(F.p))
    assert len(zs) == 3 and zs[0][0] == 1
    assert zs[-1][usable] == 1
    for a, b in zip(zs[:-1], zs[1:]):
        assert b[0] == a[usable]


@pytest.mark.parametrize("F", FIELDS, ids=["Fp", "Fq"])
def test_permute_expression_pair_properties(F):
    rng = random.Random(3)
    usable = 200
    table = [rng.randrange(40) for _ in range(usable - 5)] + [F.p - 1, F.p - 2, 1 << 200, 7, 7]
    inp = [rng.choice(table) for _ in range(usable)]
    a, s = pm.permute_expression_pair(F, inp + [99], table + [98], usable)
    assert sorted(a) == a == sorted(inp) and sorted(s) == sorted(table)
    for i in range(usable):
        assert a[i] == s[i] or (i > 0 and a[i] == a[i - 1])
    assert a[0] == s[0]
    assert pm.permute_expression_pair(F, [5, 6], [5, 5], 2) is None


@pytest.mark.parametrize("F", FIELDS, ids=["Fp", "Fq"])
def test_lookup_product_closes(F):
    rng = random.Random(4)
    k, bf = 6, 5
    n = 1 << k
    usable = n - (bf + 1)
    table = [rng.randrange(30) for _ in range(usable)]
    inp = [rng.choice(table) for _ in range(usable)]
    pa, ps = pm.permute_expression_pair(F, inp, table, usable)
    pad = lambda v: v + [rng.randrange(F.p) for _ in range(n - usable)]
    beta, gamma = rng.randrange(F.p), rng.randrange(F.p)
    z = pm.lookup_commit_product(F, n, pad(inp), pad(table), pad(pa), pad(ps), beta, gamma, bf, lambda: rng.randrange(F.p))
    assert len(z) == n and z[0] == 1 and z[usable] == 1


# ---- row f2: the oracle's opening phase ------------------------------------------------------------------------------------
@pytest.mark.parametrize("F", FIELDS, ids=["Fp", "Fq"])
def test_kate_division_identity(F):
    rng = random.Random(5)
    a = [rng.randrange(F.p) for _ in range(40)]
    for b in (rng.randrange(F.p), 0, 1):
        q = pm.kate_division(F, a, b)
        assert len(q) == 39
        r = rng.randrange(F.p)
        assert ((r - b) * pm.eval_polynomial(F, q, r) + pm.eval_polynomial(F, a, b)) % F.p == pm.eval_polynomial(F, a, r)


class _Transcript:
    def __init__(self, seed, p):
        self.rng, self.p, self.points, self.scalars, self.challenges = random.Random(seed), p, [], [], []

    def write_point(self, P): self.points.append(P)
    def write_scalar(self, s): self.scalars.append(s)

    def squeeze_challenge_scalar(self):
        c = self.rng.randrange(1, self.p)
        self.challenges.append(c)
        return c


@pytest.mark.parametrize("C", [pm.Vesta, pm.Pallas], ids=["vesta", "pallas"])
def test_ipa_proof_satisfies_the_verifier_equation(C):
    """The restated prover must produce proofs the halo2 verifier accepts (poly/commitment/verifier.rs):
    P' + sum_j [u_j^-1] L_j + [u_j] R_j = [c] G'_0 + [c * b_0 * z] U + [f] W,  P' = P - [v] G_0 + [xi] S."""
    F = C.scalar
    rng = random.Random(6)
    k = 3
    n = 1 << k
    G = (C.base.p - 1, 2)
    g = [C.mul(rng.randrange(1, F.p), G) for _ in range(n)]
    w, u = C.mul(rng.randrange(1, F.p), G), C.mul(rng.randrange(1, F.p), G)
    p_poly = [rng.randrange(F.p) for _ in range(n)]
    p_blind, x_3 = rng.randrange(F.p), rng.randrange(F.p)
    t = _Transcript(1, F.p)
    pm.ipa_create_proof(C, k, g, w, u, lambda: rng.randrange(F.p), t, p_poly, p_blind, x_3)
    s_commit, rounds = t.points[0], t.points[1:]
    xi, z, us = t.challenges[0], t.challenges[1], t.challenges[2:]
    c, f = t.scalars
    assert len(rounds) == 2 * k and len(us) == k
    P = C.best_multiexp(p_poly + [p_blind], g + [w])
    v = pm.eval_polynomial(F, p_poly, x_3)
    lhs = C.add(C.add(P, C.neg(C.mul(v, g[0]))), C.mul(xi, s_commit))
    g_fold, b = list(g), [pow(x_3, i, F.p) for i in range(n)]
    for j, u_j in enumerate(us):
        lhs = C.add(lhs, C.add(C.mul(F.inv(u_j), rounds[2 * j]), C.mul(u_j, rounds[2 * j + 1])))
        half = len(b) // 2
        b = [(b[i] + b[i + half] * u_j) % F.p for i in range(half)]
        g_fold = pm.parallel_generator_collapse(C, g_fold, u_j)
    rhs = C.add(C.add(C.mul(c, g_fold[0]), C.mul(c * b[0] % F.p * z % F.p, u)), C.mul(f, w))
    assert lhs == rhs
