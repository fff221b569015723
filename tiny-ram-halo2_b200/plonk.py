"""Mirror of halo2_proofs::plonk::{ConstraintSystem, keygen_vk, keygen_pk, create_proof} and of the Blake2b transcript
(plonk/{circuit,keygen,prover}.rs, plonk/{permutation,lookup,vanishing}/prover.rs, poly/multiopen/prover.rs, transcript.rs of
halo2_proofs 0.2.0 @ a95945254dcc, the un-vendored dependency the reference calls at /root/reference/src/test_utils.rs:21-49)
-- SURVEY.md 8(f) row f4: the proof assembly that strings the hot kernels together and produces the serialized proof.

The protocol logic below (phase order, RNG draw order, transcript order, query order, point-set construction) is host code,
as it is in the reference; every piece of field / group arithmetic goes through a BACKEND object:

    GpuBackend (this file)      the product: libtrp.so through the ctypes mirrors (commitment / domain / poly / permutation /
                                lookup / ipa).  There is no CPU fallback; constructing it without a GPU raises TrpError.
    oracle/plonk_model.py       test infrastructure only: the same interface over the Python big-int model, used by the
                                tests as the checker (identical proof bytes), next to an independent verify_proof.

Values at this level are canonical Python ints; points are affine (x, y) int tuples, None = identity.  The witness is
supplied as columns (circuit synthesis / layouting stay with the caller, as north_star keeps them on the Rust side).

Deviation that cannot be avoided here: VerifyingKey::transcript_repr hashes Rust's `{:?}` rendering of the pinned key; the same
BLAKE2b construction is applied to OUR canonical text rendering (pinned_text), so proofs are self-consistent with
oracle/plonk_model.verify_proof but the first transcript scalar differs from a Rust run's."""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

from . import poly as P

ADVICE, FIXED, INSTANCE = "advice", "fixed", "instance"


# ---- plonk::Expression ------------------------------------------------------------------------------------------------------------
class Expression:
    def __add__(self, o): return Sum(self, _lift(o))
    __radd__ = __add__
    def __neg__(self): return Negated(self)
    def __sub__(self, o): return Sum(self, Negated(_lift(o)))
    def __rsub__(self, o): return Sum(_lift(o), Negated(self))
    def __mul__(self, o): return Scaled(self, o) if isinstance(o, int) else Product(self, o)
    def __rmul__(self, o): return Scaled(self, o)


def _lift(x):
    return Constant(x) if isinstance(x, int) else x


@dataclass(eq=False)
class Constant(Expression):
    value: int
    def degree(self): return 0


@dataclass(eq=False)
class Query(Expression):
    kind: str
    column: int
    rotation: int
    def degree(self): return 1


@dataclass(eq=False)
class Negated(Expression):
    a: Expression
    def degree(self): return self.a.degree()


@dataclass(eq=False)
class Sum(Expression):
    a: Expression
    b: Expression
    def degree(self): return max(self.a.degree(), self.b.degree())


@dataclass(eq=False)
class Product(Expression):
    a: Expression
    b: Expression
    def degree(self): return self.a.degree() + self.b.degree()


@dataclass(eq=False)
class Scaled(Expression):
    a: Expression
    scalar: int
    def degree(self): return self.a.degree()


def evaluate_expression(e: Expression, p: int, query: Callable[[Query], int]) -> int:
    """Expression::evaluate over field elements (the verifier's use, and row-wise compression of lookup expressions)"""
    if isinstance(e, Constant): return e.value % p
    if isinstance(e, Query): return query(e)
    if isinstance(e, Negated): return -evaluate_expression(e.a, p, query) % p
    if isinstance(e, Sum): return (evaluate_expression(e.a, p, query) + evaluate_expression(e.b, p, query)) % p
    if isinstance(e, Product): return evaluate_expression(e.a, p, query) * evaluate_expression(e.b, p, query) % p
    if isinstance(e, Scaled): return evaluate_expression(e.a, p, query) * e.scalar % p
    raise TypeError(type(e))


# ---- plonk::ConstraintSystem ------------------------------------------------------------------------------------------------------
class ConstraintSystem:
    """The subset of plonk::ConstraintSystem the prover reads: columns, ordered queries, gate polynomials, lookups, the
    permutation's columns.  Selectors are plain fixed columns here (halo2 compresses them into fixed columns at keygen)."""

    def __init__(self):
        self.num_advice = self.num_fixed = self.num_instance = 0
        self.queries = {ADVICE: [], FIXED: [], INSTANCE: []}     # ordered (column, rotation), as cs.*_queries
        self.num_advice_queries: List[int] = []
        self.gates: List[Expression] = []
        self.lookups: List[Tuple[List[Expression], List[Expression]]] = []
        self.permutation: List[Tuple[str, int]] = []
        self.minimum_degree: Optional[int] = None

    def advice_column(self):
        self.num_advice += 1; self.num_advice_queries.append(0); return self.num_advice - 1

    def fixed_column(self):
        self.num_fixed += 1; return self.num_fixed - 1

    def instance_column(self):
        self.num_instance += 1; return self.num_instance - 1

    def query(self, kind: str, column: int, rotation: int = 0) -> Query:
        """meta.query_advice / query_fixed / query_instance: registers (column, rotation) on first use"""
        if (column, rotation) not in self.queries[kind]:
            self.queries[kind].append((column, rotation))
            if kind == ADVICE:
                self.num_advice_queries[column] += 1
        return Query(kind, column, rotation)

    def enable_equality(self, kind: str, column: int):
        self.query(kind, column, 0)
        if (kind, column) not in self.permutation:
            self.permutation.append((kind, column))

    def create_gate(self, polys: Sequence[Expression]):
        self.gates.extend(polys)

    def lookup(self, pairs: Sequence[Tuple[Expression, Expression]]):
        self.lookups.append(([a for a, _ in pairs], [t for _, t in pairs]))

    def query_index(self, kind, column, rotation):
        return self.queries[kind].index((column, rotation))

    def degree(self) -> int:
        degree = 3                                              # permutation::Argument::required_degree
        for inputs, tables in self.lookups:
            di = max([1] + [e.degree() for e in inputs]); dt = max([1] + [e.degree() for e in tables])
            degree = max(degree, max(4, 2 + di + dt))           # lookup::Argument::required_degree
        degree = max([degree] + [g.degree() for g in self.gates])
        return max(degree, self.minimum_degree or 1)

    def blinding_factors(self) -> int:
        factors = max(self.num_advice_queries + [1])
        return max(3, factors) + 2

    def minimum_rows(self) -> int:
        return self.blinding_factors() + 3

    def pinned_text(self) -> str:
        def ex(e):
            if isinstance(e, Constant): return f"Constant({e.value:#x})"
            if isinstance(e, Query): return f"{e.kind.capitalize()}({e.column}, {e.rotation})"
            if isinstance(e, Negated): return f"Negated({ex(e.a)})"
            if isinstance(e, Sum): return f"Sum({ex(e.a)}, {ex(e.b)})"
            if isinstance(e, Product): return f"Product({ex(e.a)}, {ex(e.b)})"
            return f"Scaled({ex(e.a)}, {e.scalar:#x})"
        return (f"PinnedConstraintSystem {{ num_fixed_columns: {self.num_fixed}, num_advice_columns: {self.num_advice}, "
                f"num_instance_columns: {self.num_instance}, gates: [{', '.join(ex(g) for g in self.gates)}], "
                f"advice_queries: {self.queries[ADVICE]}, instance_queries: {self.queries[INSTANCE]}, fixed_queries: {self.queries[FIXED]}, "
                f"permutation: {self.permutation}, lookups: [{', '.join('(' + ', '.join(ex(e) for e in a) + ' -> ' + ', '.join(ex(e) for e in t) + ')' for a, t in self.lookups)}], "
                f"minimum_degree: {self.minimum_degree} }}")


# ---- transcript.rs: Blake2bWrite<_, _, Challenge255<_>> ---------------------------------------------------------------------------
class Blake2bWrite:
    """BLAKE2b-512 personalised "Halo2-Transcript"; prefixes 0 (challenge), 1 (point), 2 (scalar); points enter the hash as
    x || y (32-byte little-endian each) and the proof as the 32-byte compressed encoding; a challenge is the 64-byte digest
    read little-endian and reduced (Challenge255 / from_bytes_wide)."""

    def __init__(self, base_modulus: int, scalar_modulus: int):
        self.state = hashlib.blake2b(digest_size=64, person=b"Halo2-Transcript")
        self.q, self.p = base_modulus, scalar_modulus
        self.proof = bytearray()

    def common_point(self, pt):
        if pt is None:
            raise ValueError("cannot write points at infinity to the transcript")
        self.state.update(b"\x01" + pt[0].to_bytes(32, "little") + pt[1].to_bytes(32, "little"))

    def common_scalar(self, s):
        self.state.update(b"\x02" + (s % self.p).to_bytes(32, "little"))

    def write_point(self, pt):
        self.common_point(pt)
        b = bytearray(pt[0].to_bytes(32, "little"))
        b[31] |= (pt[1] & 1) << 7
        self.proof += b

    def write_scalar(self, s):
        self.common_scalar(s)
        self.proof += (s % self.p).to_bytes(32, "little")

    def squeeze_challenge_scalar(self) -> int:
        self.state.update(b"\x00")
        return int.from_bytes(self.state.copy().digest(), "little") % self.p

    def finalize(self) -> bytes:
        return bytes(self.proof)


def transcript_repr(text: str, p: int) -> int:
    """VerifyingKey::from_parts: BLAKE2b-512 personalised "Halo2-Verify-Key" over len(text) (u64 LE) || text, from_bytes_wide"""
    h = hashlib.blake2b(digest_size=64, person=b"Halo2-Verify-Key")
    h.update(len(text).to_bytes(8, "little") + text.encode())
    return int.from_bytes(h.digest(), "little") % p


# ---- keys -------------------------------------------------------------------------------------------------------------------------
@dataclass
class VerifyingKey:
    k: int
    cs: ConstraintSystem
    cs_degree: int
    fixed_commitments: list
    permutation_commitments: list
    transcript_repr: int


@dataclass
class ProvingKey:
    vk: VerifyingKey
    fixed_values: list
    fixed_polys: list
    fixed_cosets: list
    sigma_values: list
    sigma_polys: list
    sigma_cosets: list
    l0: object
    l_blind: object
    l_last: object


def build_sigmas(cs: ConstraintSystem, n: int, p: int, omega: int, delta: int, copies):
    """permutation::keygen::Assembly (copy -> cycle merge) and build_pk's sigma values: sigma_i[j] = delta^i' * omega^j' for the
    cell (i', j') that (i, j) maps to.  copies: ((kind, column, row), (kind, column, row)) pairs."""
    m = len(cs.permutation)
    col_of = {kc: i for i, kc in enumerate(cs.permutation)}
    mapping = [[(i, j) for j in range(n)] for i in range(m)]
    aux = [[(i, j) for j in range(n)] for i in range(m)]
    sizes = [[1] * n for _ in range(m)]
    for (lk, lc, lr), (rk, rc, rr) in copies:
        if (lk, lc) not in col_of or (rk, rc) not in col_of:
            raise ValueError("copy constraint on a column without enable_equality")
        left, right = (col_of[(lk, lc)], lr), (col_of[(rk, rc)], rr)
        lcyc, rcyc = aux[left[0]][left[1]], aux[right[0]][right[1]]
        if lcyc == rcyc:
            continue
        if sizes[lcyc[0]][lcyc[1]] < sizes[rcyc[0]][rcyc[1]]:
            lcyc, rcyc = rcyc, lcyc
        sizes[lcyc[0]][lcyc[1]] += sizes[rcyc[0]][rcyc[1]]
        i = rcyc
        while True:
            aux[i[0]][i[1]] = lcyc
            i = mapping[i[0]][i[1]]
            if i == rcyc:
                break
        mapping[left[0]][left[1]], mapping[right[0]][right[1]] = mapping[right[0]][right[1]], mapping[left[0]][left[1]]
    om = [1] * n
    for j in range(1, n):
        om[j] = om[j - 1] * omega % p
    dl = [1] * max(m, 1)
    for i in range(1, m):
        dl[i] = dl[i - 1] * delta % p
    return [[dl[mapping[i][j][0]] * om[mapping[i][j][1]] % p for j in range(n)] for i in range(m)]


def keygen(backend, cs: ConstraintSystem, fixed_values, copies=()) -> ProvingKey:
    """keygen_vk + keygen_pk.  fixed_values: one list per fixed column (padded with zeros to n)."""
    n, p = backend.n, backend.p
    if n < cs.minimum_rows():
        raise ValueError("NotEnoughRowsAvailable")
    if backend.j != cs.degree():
        raise ValueError("the backend's EvaluationDomain was built for a different constraint-system degree")
    pad = lambda v: [x % p for x in v] + [0] * (n - len(v))
    fixed_values = [pad(v) for v in fixed_values]
    if len(fixed_values) != cs.num_fixed or any(len(v) != n for v in fixed_values):
        raise ValueError("one column of at most n values per fixed column")
    fixed_commitments = [backend.commit_lagrange(v, 1) for v in fixed_values]          # Blind::default() = 1
    sigma_values = build_sigmas(cs, n, p, backend.omega, backend.delta, copies)
    permutation_commitments = [backend.commit_lagrange(v, 1) for v in sigma_values]
    pt = lambda c: "Identity" if c is None else f"({c[0]:#066x}, {c[1]:#066x})"
    text = (f"PinnedVerificationKey {{ base_modulus: {backend.q:#066x}, scalar_modulus: {p:#066x}, domain: PinnedEvaluationDomain {{ k: {backend.k}, "
            f"extended_k: {backend.extended_k}, omega: {backend.omega:#066x} }}, cs: {cs.pinned_text()}, "
            f"fixed_commitments: [{', '.join(pt(c) for c in fixed_commitments)}], permutation: VerifyingKey {{ commitments: "
            f"[{', '.join(pt(c) for c in permutation_commitments)}] }} }}")
    vk = VerifyingKey(backend.k, cs, cs.degree(), fixed_commitments, permutation_commitments, transcript_repr(text, p))
    fixed_polys = [backend.lagrange_to_coeff(v) for v in fixed_values]
    sigma_polys = [backend.lagrange_to_coeff(v) for v in sigma_values]
    bf = cs.blinding_factors()
    ext = lambda lag: backend.coeff_to_extended(backend.lagrange_to_coeff(lag))
    l0 = [0] * n; l0[0] = 1
    l_blind = [0] * (n - bf) + [1] * bf
    l_last = [0] * n; l_last[n - bf - 1] = 1
    return ProvingKey(vk, fixed_values, fixed_polys, [backend.coeff_to_extended(c) for c in fixed_polys], sigma_values, sigma_polys,
                      [backend.coeff_to_extended(c) for c in sigma_polys], ext(l0), ext(l_blind), ext(l_last))


# ---- poly/multiopen: construct_intermediate_sets -----------------------------------------------------------------------------------
def construct_intermediate_sets(queries, commitment_key: Callable, point_of: Callable, eval_of: Callable):
    """poly::multiopen::construct_intermediate_sets.  Returns (commitment_map, point_sets): commitment_map is a list of dicts
    {commitment, set_index, point_indices, evals} in order of first appearance; point_sets[set] lists that set's points."""
    point_index_map = {}
    for q in queries:
        point_index_map.setdefault(point_of(q), len(point_index_map))
    inverse = {v: k for k, v in point_index_map.items()}
    cmap, where = [], {}
    for q in queries:
        key = commitment_key(q)
        if key not in where:
            where[key] = len(cmap)
            cmap.append({"commitment": q, "point_indices": [], "set_index": None, "evals": None})
        pi = point_index_map[point_of(q)]
        if pi not in cmap[where[key]]["point_indices"]:
            cmap[where[key]]["point_indices"].append(pi)
    point_idx_sets = {}
    for cd in cmap:
        point_idx_sets.setdefault(tuple(sorted(cd["point_indices"])), len(point_idx_sets))
    for cd in cmap:
        cd["evals"] = [None] * len(cd["point_indices"])
    for q in queries:
        cd = cmap[where[commitment_key(q)]]
        pset = tuple(sorted(cd["point_indices"]))
        cd["set_index"] = point_idx_sets[pset]
        cd["evals"][pset.index(point_index_map[point_of(q)])] = eval_of(q)
    point_sets = [None] * len(point_idx_sets)
    for pset, idx in point_idx_sets.items():
        point_sets[idx] = [inverse[i] for i in pset]
    return cmap, point_sets


def lagrange_interpolate(points, evals, p):
    """arithmetic::lagrange_interpolate: coefficients (low -> high) of the polynomial of degree < len(points) through the pairs"""
    m = len(points)
    if m == 1:
        return [evals[0] % p]
    out = [0] * m
    for j in range(m):
        num, den = [1], 1
        for i in range(m):
            if i == j:
                continue
            nxt = [0] * (len(num) + 1)
            for d, c in enumerate(num):
                nxt[d] = (nxt[d] - c * points[i]) % p
                nxt[d + 1] = (nxt[d + 1] + c) % p
            num = nxt
            den = den * (points[j] - points[i]) % p
        s = evals[j] * pow(den, -1, p) % p
        for d, c in enumerate(num):
            out[d] = (out[d] + c * s) % p
    return out


# ---- plonk::create_proof -----------------------------------------------------------------------------------------------------------
@dataclass
class _Opening:
    poly: list          # coefficient form
    blind: int
    point: int
    eval: int = 0


def create_proof(backend, pk: ProvingKey, instances, advice, rand: Callable[[], int], transcript: Blake2bWrite) -> bytes:
    """plonk::create_proof for ONE circuit instance.  instances: one list per instance column (at most usable_rows values);
    advice: one list per advice column (values on the usable rows; the blinding rows are drawn here).  rand() draws one
    uniformly random scalar (the caller's RNG: every draw happens in halo2's order).  Returns the proof bytes (also left in
    `transcript`)."""
    vk, cs = pk.vk, pk.vk.cs
    n, p, k = backend.n, backend.p, backend.k
    bf = cs.blinding_factors()
    usable = n - (bf + 1)
    rot = backend.rotate_omega
    if len(instances) != cs.num_instance or len(advice) != cs.num_advice:
        raise ValueError("InvalidInstances / wrong number of advice columns")
    transcript.common_scalar(vk.transcript_repr)

    # ---- instance columns: commit (not written, only absorbed) -------------------------------------------------------------
    inst_values = []
    for col in instances:
        if len(col) > usable:
            raise ValueError("InstanceTooLarge")
        inst_values.append([v % p for v in col] + [0] * (n - len(col)))
    for v in inst_values:
        transcript.common_point(backend.commit_lagrange(v, 1))
    inst_polys = [backend.lagrange_to_coeff(v) for v in inst_values]
    inst_cosets = [backend.coeff_to_extended(c) for c in inst_polys]

    # ---- advice columns --------------------------------------------------------------------------------------------------------
    adv_values = []
    for col in advice:
        if len(col) > usable:
            raise ValueError("advice column longer than the usable rows")
        adv_values.append([v % p for v in col] + [0] * (n - len(col)))
    for v in adv_values:
        for r in range(usable, n):
            v[r] = rand()
    adv_blinds = [rand() for _ in adv_values]
    for v, b in zip(adv_values, adv_blinds):
        transcript.write_point(backend.commit_lagrange(v, b))
    adv_polys = [backend.lagrange_to_coeff(v) for v in adv_values]
    adv_cosets = [backend.coeff_to_extended(c) for c in adv_polys]
    values_of = {ADVICE: adv_values, FIXED: pk.fixed_values, INSTANCE: inst_values}

    # ---- lookups: commit_permuted -------------------------------------------------------------------------------------------------
    theta = transcript.squeeze_challenge_scalar()

    def compress(exprs):
        out = [0] * n
        for e in exprs:
            for r in range(n):
                out[r] = (out[r] * theta + evaluate_expression(e, p, lambda q: values_of[q.kind][q.column][(r + q.rotation) % n])) % p
        return out

    lookups = []
    for inputs, tables in cs.lookups:
        ci, ct = compress(inputs), compress(tables)
        pi, pt = backend.permute_expression_pair(ci, ct, usable)
        pi = list(pi) + [rand() for _ in range(bf + 1)]
        pt = list(pt) + [rand() for _ in range(bf + 1)]
        L = {"ci": ci, "ct": ct, "pi": pi, "pt": pt}
        for name in ("pi", "pt"):
            L[name + "_poly"] = backend.lagrange_to_coeff(L[name])
            L[name + "_blind"] = rand()
            transcript.write_point(backend.commit_lagrange(L[name], L[name + "_blind"]))
        lookups.append(L)

    # ---- permutation and lookup grand products ----------------------------------------------------------------------------------------
    beta = transcript.squeeze_challenge_scalar()
    gamma = transcript.squeeze_challenge_scalar()
    chunk_len = vk.cs_degree - 2
    perm_sets = []

    def after_chunk(z):
        blind = rand()
        transcript.write_point(backend.commit_lagrange(z, blind))
        perm_sets.append({"z": z, "blind": blind})

    if cs.permutation:
        backend.permutation_commit([values_of[kk][c] for kk, c in cs.permutation], pk.sigma_values, beta, gamma, chunk_len, bf, rand, after_chunk)
    for S in perm_sets:
        S["poly"] = backend.lagrange_to_coeff(S["z"])
        S["coset"] = backend.coeff_to_extended(S["poly"])
    for L in lookups:
        L["z"] = backend.lookup_product(L["ci"], L["ct"], L["pi"], L["pt"], beta, gamma, bf, rand)
        L["z_blind"] = rand()
        transcript.write_point(backend.commit_lagrange(L["z"], L["z_blind"]))
        L["z_poly"] = backend.lagrange_to_coeff(L["z"])

    # ---- vanishing argument: random polynomial ------------------------------------------------------------------------------------------
    random_poly = [rand() for _ in range(n)]
    random_blind = rand()
    transcript.write_point(backend.commit(random_poly, random_blind))

    # ---- quotient --------------------------------------------------------------------------------------------------------------------
    y = transcript.squeeze_challenge_scalar()
    leaves, ext_polys = {}, []

    def leaf(key, coset):
        if key not in leaves:
            leaves[key] = len(ext_polys)
            ext_polys.append(coset)
        return P.Poly(leaves[key], 0)

    cos_of = {ADVICE: adv_cosets, FIXED: pk.fixed_cosets, INSTANCE: inst_cosets}

    def to_ast(e):
        if isinstance(e, Constant): return P.ConstantTerm(e.value % p)
        if isinstance(e, Query): return leaf((e.kind, e.column), cos_of[e.kind][e.column]).with_rotation(e.rotation)
        if isinstance(e, Negated): return -to_ast(e.a)
        if isinstance(e, Sum): return to_ast(e.a) + to_ast(e.b)
        if isinstance(e, Product): return to_ast(e.a) * to_ast(e.b)
        return to_ast(e.a) * (e.scalar % p)

    l0, l_blind, l_last = leaf("l0", pk.l0), leaf("l_blind", pk.l_blind), leaf("l_last", pk.l_last)
    one = P.ConstantTerm(1)
    active = one - (l_last + l_blind)
    exprs = [to_ast(g) for g in cs.gates]
    if perm_sets:
        zs = [leaf(("perm_z", i), S["coset"]) for i, S in enumerate(perm_sets)]
        exprs.append(l0 * (one - zs[0]))
        exprs.append(l_last * (zs[-1] * zs[-1] - zs[-1]))
        for i in range(1, len(zs)):
            exprs.append(l0 * (zs[i] - zs[i - 1].with_rotation(-(bf + 1))))
        for i, z in enumerate(zs):
            cols = cs.permutation[i * chunk_len:(i + 1) * chunk_len]
            left, right = z.with_rotation(1), z
            cur_delta = beta * pow(backend.delta, i * chunk_len, p) % p
            for off, (kk, c) in enumerate(cols):
                col = leaf((kk, c), cos_of[kk][c])
                sig = leaf(("sigma", i * chunk_len + off), pk.sigma_cosets[i * chunk_len + off])
                left = left * (col + sig * beta + gamma)
                right = right * (col + P.LinearTerm(cur_delta) + gamma)
                cur_delta = cur_delta * backend.delta % p
            exprs.append((left - right) * active)
    for li, (L, (inputs, tables)) in enumerate(zip(lookups, cs.lookups)):
        z = leaf(("lookup_z", li), backend.coeff_to_extended(L["z_poly"]))
        a = leaf(("lookup_a", li), backend.coeff_to_extended(L["pi_poly"]))
        s = leaf(("lookup_s", li), backend.coeff_to_extended(L["pt_poly"]))
        comp = lambda es: P.DistributePowers([to_ast(e) for e in es], P.ConstantTerm(theta)) if len(es) > 1 else to_ast(es[0])
        exprs.append(l0 * (one - z))
        exprs.append(l_last * (z * z - z))
        exprs.append((z.with_rotation(1) * (a + beta) * (s + gamma) - z * (comp(inputs) + beta) * (comp(tables) + gamma)) * active)
        exprs.append(l0 * (a - s))
        exprs.append(((a - s) * (a - a.with_rotation(-1))) * active)
    h_ast = P.ConstantTerm(0)
    for e in exprs:
        h_ast = h_ast * y + e
    h_coeffs = backend.quotient(h_ast, ext_polys)                       # n * (j - 1) coefficients of h(X)
    pieces = [h_coeffs[i * n:(i + 1) * n] for i in range(vk.cs_degree - 1)]
    h_blinds = [rand() for _ in pieces]
    for piece, b in zip(pieces, h_blinds):
        transcript.write_point(backend.commit(piece, b))

    # ---- evaluations -----------------------------------------------------------------------------------------------------------------
    x = transcript.squeeze_challenge_scalar()
    xn = pow(x, n, p)
    polys_of = {ADVICE: adv_polys, FIXED: pk.fixed_polys, INSTANCE: inst_polys}
    for kind in (INSTANCE, ADVICE, FIXED):
        for c, r in cs.queries[kind]:
            transcript.write_scalar(backend.eval_polynomial(polys_of[kind][c], rot(x, r)))
    h_poly, h_blind = [0] * n, 0
    for piece, b in zip(reversed(pieces), reversed(h_blinds)):
        h_poly = [(a * xn + c) % p for a, c in zip(h_poly, piece)]
        h_blind = (h_blind * xn + b) % p
    transcript.write_scalar(backend.eval_polynomial(random_poly, x))
    for sp in pk.sigma_polys:
        transcript.write_scalar(backend.eval_polynomial(sp, x))
    x_next, x_inv, x_last = rot(x, 1), rot(x, -1), rot(x, -(bf + 1))
    for i, S in enumerate(perm_sets):
        transcript.write_scalar(backend.eval_polynomial(S["poly"], x))
        transcript.write_scalar(backend.eval_polynomial(S["poly"], x_next))
        if i + 1 < len(perm_sets):
            transcript.write_scalar(backend.eval_polynomial(S["poly"], x_last))
    for L in lookups:
        for poly, pt_ in ((L["z_poly"], x), (L["z_poly"], x_next), (L["pi_poly"], x), (L["pi_poly"], x_inv), (L["pt_poly"], x)):
            transcript.write_scalar(backend.eval_polynomial(poly, pt_))

    # ---- multiopen ---------------------------------------------------------------------------------------------------------------------
    qs: List[_Opening] = []
    for c, r in cs.queries[INSTANCE]:
        qs.append(_Opening(inst_polys[c], 1, rot(x, r)))
    for c, r in cs.queries[ADVICE]:
        qs.append(_Opening(adv_polys[c], adv_blinds[c], rot(x, r)))
    for S in perm_sets:
        qs.append(_Opening(S["poly"], S["blind"], x))
        qs.append(_Opening(S["poly"], S["blind"], x_next))
    for S in list(reversed(perm_sets))[1:]:
        qs.append(_Opening(S["poly"], S["blind"], x_last))
    for L in lookups:
        qs.append(_Opening(L["z_poly"], L["z_blind"], x))
        qs.append(_Opening(L["pi_poly"], L["pi_blind"], x))
        qs.append(_Opening(L["pt_poly"], L["pt_blind"], x))
        qs.append(_Opening(L["pi_poly"], L["pi_blind"], x_inv))
        qs.append(_Opening(L["z_poly"], L["z_blind"], x_next))
    for c, r in cs.queries[FIXED]:
        qs.append(_Opening(pk.fixed_polys[c], 1, rot(x, r)))
    for sp in pk.sigma_polys:
        qs.append(_Opening(sp, 1, x))
    qs.append(_Opening(h_poly, h_blind, x))
    qs.append(_Opening(random_poly, random_blind, x))
    for q in qs:
        q.eval = backend.eval_polynomial(q.poly, q.point)

    x_1 = transcript.squeeze_challenge_scalar()
    x_2 = transcript.squeeze_challenge_scalar()
    cmap, point_sets = construct_intermediate_sets(qs, lambda q: (id(q.poly), q.blind), lambda q: q.point, lambda q: q.eval)
    ns = len(point_sets)
    q_polys, q_blinds = [None] * ns, [0] * ns
    q_eval_sets = [[0] * len(ps) for ps in point_sets]
    for cd in cmap:
        s, o = cd["set_index"], cd["commitment"]
        q_polys[s] = list(o.poly) if q_polys[s] is None else [(a * x_1 + c) % p for a, c in zip(q_polys[s], o.poly)]
        q_blinds[s] = (q_blinds[s] * x_1 + o.blind) % p
        q_eval_sets[s] = [(a * x_1 + e) % p for a, e in zip(q_eval_sets[s], cd["evals"])]
    q_prime = None
    for points, evals, poly in zip(point_sets, q_eval_sets, q_polys):
        r_poly = lagrange_interpolate(points, evals, p)
        cur = list(poly)
        for i, r in enumerate(r_poly):
            cur[i] = (cur[i] - r) % p
        for pt_ in points:
            cur = backend.kate_division(cur, pt_)
        cur = list(cur) + [0] * (n - len(cur))
        q_prime = cur if q_prime is None else [(a * x_2 + c) % p for a, c in zip(q_prime, cur)]
    q_prime_blind = rand()
    transcript.write_point(backend.commit(q_prime, q_prime_blind))
    x_3 = transcript.squeeze_challenge_scalar()
    for qp in q_polys:
        transcript.write_scalar(backend.eval_polynomial(qp, x_3))
    x_4 = transcript.squeeze_challenge_scalar()
    p_poly, p_blind = q_prime, q_prime_blind
    for qp, qb in zip(q_polys, q_blinds):
        p_poly = [(a * x_4 + c) % p for a, c in zip(p_poly, qp)]
        p_blind = (p_blind * x_4 + qb) % p
    backend.ipa_create_proof(rand, transcript, p_poly, p_blind, x_3)
    return transcript.finalize()


# ---- the product backend: libtrp.so ---------------------------------------------------------------------------------------------------
class GpuBackend:
    """Backend of create_proof / keygen over the CUDA library (host-array entry points of include/tr_prover.h)."""

    def __init__(self, ctx, k: int, cs_degree: int, params=None):
        import numpy as np
        from . import ipa as _ipa, lookup as _lookup, permutation as _perm
        from .commitment import Params
        from .domain import EvaluationDomain
        self.np, self._ipa, self._lookup, self._perm = np, _ipa, _lookup, _perm
        self.ctx, self.k, self.n, self.j = ctx, k, 1 << k, cs_degree
        self.p = _perm._MODULUS[ctx.curve]
        self.q = _perm._MODULUS[1 - ctx.curve]
        self.R, self.Rq = (1 << 256) % self.p, (1 << 256) % self.q
        self.Rinv, self.Rqinv = pow(self.R, -1, self.p), pow(self.Rq, -1, self.q)
        self.params = params if params is not None else Params.new(ctx, k)
        self.dom = EvaluationDomain(ctx, cs_degree, k)
        self.extended_k = self.dom.extended_k
        self.omega = self._ints(self.dom.omega.reshape(1, 4))[0]
        self.omega_inv = pow(self.omega, -1, self.p)
        self.delta = pow(5, 1 << 32, self.p)
        self.ipa_params = _ipa.IpaParams(ctx, k, self.params.g_points, self.params.w, self.params.u)

    # -- conversions between canonical ints and Montgomery limb arrays
    def _limbs(self, vals, mod=None, R=None):
        np = self.np
        mod, R = mod or self.p, R or self.R
        buf = b"".join((v % mod * R % mod).to_bytes(32, "little") for v in vals)
        return np.frombuffer(buf, dtype=np.uint64).reshape(len(vals), 4).copy()

    def _ints(self, arr, mod=None, Rinv=None):
        mod, Rinv = mod or self.p, Rinv or self.Rinv
        raw = self.np.ascontiguousarray(arr, dtype=self.np.uint64).reshape(-1, 4).tobytes()
        return [int.from_bytes(raw[i:i + 32], "little") * Rinv % mod for i in range(0, len(raw), 32)]

    def _point(self, jac):
        """normalised Jacobian (3, 4) -> affine tuple / None"""
        if not jac[2].any():
            return None
        x, y = self._ints(jac[:2], self.q, self.Rqinv)
        return (x, y)

    def rotate_omega(self, x, rotation):
        return x * pow(self.omega if rotation >= 0 else self.omega_inv, abs(rotation), self.p) % self.p

    def commit_lagrange(self, values, blind):
        return self._point(self.params.commit_lagrange(self._limbs(values), self._limbs([blind])[0]))

    def commit(self, coeffs, blind):
        return self._point(self.params.commit(self._limbs(coeffs), self._limbs([blind])[0]))

    def lagrange_to_coeff(self, values):
        return self._ints(self.dom.lagrange_to_coeff(self._limbs(values)))

    def coeff_to_extended(self, coeffs):
        return self.dom.coeff_to_extended(self._limbs(coeffs))            # opaque: (2^extended_k, 4) Montgomery array

    def quotient(self, ast, ext_polys):
        ev = P.new_evaluator(self.ctx)
        for e in ext_polys:
            ev.register_poly(e)
        h_ext = ev.evaluate(ast, self.dom)
        return self._ints(self.dom.extended_to_coeff(h_ext, divide_by_vanishing_poly=True))

    def eval_polynomial(self, coeffs, x):
        return self._ints(self._ipa.eval_polynomial(self.ctx, self._limbs(coeffs), self._limbs([x])[0]).reshape(1, 4))[0]

    def kate_division(self, coeffs, b):
        return self._ints(self._ipa.kate_division(self.ctx, self._limbs(coeffs), self._limbs([b])[0]))

    def permutation_commit(self, values, sigmas, beta, gamma, chunk_len, blinding_factors, rand, after_chunk):
        np = self.np
        v = np.stack([self._limbs(c) for c in values]); s = np.stack([self._limbs(c) for c in sigmas])
        return self._perm.commit(self.dom, v, s, beta, gamma, chunk_len, blinding_factors, rand, lambda z: after_chunk(self._ints(z)))

    def permute_expression_pair(self, inp, tab, usable_rows):
        a, s = self._lookup.permute_expression_pair(self.ctx, self._limbs(inp), self._limbs(tab), usable_rows)
        return self._ints(a), self._ints(s)

    def lookup_product(self, ci, ct, pi, pt, beta, gamma, blinding_factors, rand):
        return self._ints(self._lookup.commit_product(self.dom, self._limbs(ci), self._limbs(ct), self._limbs(pi), self._limbs(pt),
                                                      beta, gamma, blinding_factors, rand))

    def ipa_create_proof(self, rand, transcript, p_poly, p_blind, x_3):
        outer = self

        class _Adapter:
            def write_point(self, limbs):
                x, y = outer._ints(outer.np.asarray(limbs, dtype=outer.np.uint64).reshape(2, 4), outer.q, outer.Rqinv)
                transcript.write_point(None if (x, y) == (0, 0) else (x, y))
            def write_scalar(self, s): transcript.write_scalar(s)
            def squeeze_challenge_scalar(self): return transcript.squeeze_challenge_scalar()

        self._ipa.create_proof(self.ipa_params, rand, _Adapter(), self._limbs(p_poly), p_blind, x_3)
