"""Mirror of halo2_proofs::plonk::{verify_proof, SingleVerifier, BatchVerifier}, poly::multiopen::verify_proof,
poly::commitment::{MSM, Guard, verify_proof} and transcript::Blake2bRead (plonk/verifier.rs, plonk/verifier/batch.rs,
poly/multiopen/verifier.rs, poly/commitment/{msm,verifier}.rs, transcript.rs of halo2_proofs 0.2.0 @ a95945254dcc, the
un-vendored dependency the reference calls at /root/reference/src/test_utils.rs:51-70 and 104-118) -- the second half of
SURVEY.md 8(f) row f4: `gen_proofs_and_verify` checks every proof it makes, so a drop-in for the prover path has to be able to
check its own output without the test oracle.

As in plonk.py, the protocol is host logic and every piece of group / vector arithmetic goes through the BACKEND
(plonk.GpuBackend in the product; oracle/plonk_model.PythonBackend in the CPU tests).  Like halo2's verifier nothing is
multiplied on the way: every commitment the check touches is appended to ONE multi-scalar multiplication

    P - [v] G_0 + [xi] S + sum_j ([u_j^-1] L_j + [u_j] R_j) - [c] <s, G> - [c b z] U - [f] W  ==  identity

whose n-term part <s, G> runs over the resident generator table (backend.commit) and whose few hundred other terms go out as
one variable-base MSM (backend.msm_points).  The backend supplies three methods beyond the prover's:

    fixed_points()          -> (G_0, W, U) as affine int tuples
    ipa_s_vector(us, init)  -> backend vector s with s_i = init * prod_j u_j^(bit (k-1-j) of i)   (commitment::compute_s)
    msm_points(scalars, points) -> affine int tuple or None (identity)

oracle/plonk_model.verify_proof stays the tests' INDEPENDENT checker; the CPU tests hold the two verifiers against each other
on accepted, tampered and malformed proofs (tests/test_verifier_cpu.py)."""
from __future__ import annotations

import hashlib
import os
from typing import Callable, List, Optional

from .plonk import ADVICE, FIXED, INSTANCE, construct_intermediate_sets, evaluate_expression, lagrange_interpolate

CURVE_B = 5                       # y^2 = x^3 + 5 on both Pasta curves


class VerifyError(Exception):
    """plonk::Error on the verifier's side (ConstraintSystemFailure, InvalidInstances, InstanceTooLarge, Opening, Transcript)"""


# ---- transcript::Blake2bRead ------------------------------------------------------------------------------------------------------------
_TS = {}                        # per modulus: (s, t, z^t) with q - 1 = 2^s t, z a non-residue


def _sqrt(a: int, q: int) -> Optional[int]:
    """a square root of a modulo the prime q (Tonelli-Shanks; the Pasta primes have 2-adicity 32), None for non-residues.
    One full-size exponentiation per call: w = a^((t-1)/2) gives a^((t+1)/2) = a w and a^t = a w^2; a non-residue shows as an
    element of full order 2^s in the loop."""
    a %= q
    if a == 0:
        return 0
    if q not in _TS:
        s, t = 0, q - 1
        while t % 2 == 0:
            s, t = s + 1, t // 2
        z = 2
        while pow(z, (q - 1) // 2, q) != q - 1:
            z += 1
        _TS[q] = (s, t, pow(z, t, q))
    m, t, c = _TS[q]
    w = pow(a, (t - 1) // 2, q)
    r = a * w % q
    u = r * w % q
    while u != 1:
        i, v = 0, u
        while v != 1:
            v, i = v * v % q, i + 1
            if i == m:
                return None                      # u has order 2^m: a is not a square
        b = c
        for _ in range(m - i - 1):
            b = b * b % q
        m, c, u, r = i, b * b % q, u * b * b % q, r * b % q
    return r


class Blake2bRead:
    """The reading side of plonk.Blake2bWrite: same hash state, points and scalars come out of the proof bytes.  A point is 32
    bytes: x little-endian in the low 255 bits, the parity of y in the top bit (pasta_curves' GroupEncoding); the identity, a
    non-canonical x and an x that is not on the curve are rejected (the transcript cannot absorb the identity)."""

    def __init__(self, proof: bytes, base_modulus: int, scalar_modulus: int):
        self.state = hashlib.blake2b(digest_size=64, person=b"Halo2-Transcript")
        self.buf, self.pos = bytes(proof), 0
        self.q, self.p = base_modulus, scalar_modulus

    def _take(self, count: int) -> bytes:
        if self.pos + count > len(self.buf):
            raise VerifyError("Transcript: the proof ends early")
        self.pos += count
        return self.buf[self.pos - count:self.pos]

    def common_point(self, pt):
        if pt is None:
            raise VerifyError("Transcript: cannot write points at infinity to the transcript")
        self.state.update(b"\x01" + pt[0].to_bytes(32, "little") + pt[1].to_bytes(32, "little"))

    def common_scalar(self, s: int):
        self.state.update(b"\x02" + (s % self.p).to_bytes(32, "little"))

    def read_point(self):
        raw = self._take(32)
        sign = raw[31] >> 7
        x = int.from_bytes(raw, "little") & ((1 << 255) - 1)
        if x >= self.q:
            raise VerifyError("Transcript: invalid point encoding in proof")
        if x == 0 and sign == 0:
            raise VerifyError("Transcript: cannot write points at infinity to the transcript")
        y = _sqrt(x * x * x + CURVE_B, self.q)
        if y is None:
            raise VerifyError("Transcript: invalid point encoding in proof")
        if y & 1 != sign:
            y = self.q - y
        self.common_point((x, y))
        return (x, y)

    def read_scalar(self) -> int:
        s = int.from_bytes(self._take(32), "little")
        if s >= self.p:
            raise VerifyError("Transcript: invalid field element encoding in proof")
        self.common_scalar(s)
        return s

    def squeeze_challenge_scalar(self) -> int:
        self.state.update(b"\x00")
        return int.from_bytes(self.state.copy().digest(), "little") % self.p


# ---- poly::commitment::MSM -----------------------------------------------------------------------------------------------------------
class MSM:
    """A multi-scalar multiplication collected lazily: terms over arbitrary points, a scalar vector over the generators g
    (backend vector, None = zero), an extra coefficient of G_0, and the coefficients of W and U."""

    def __init__(self, backend):
        self.B = backend
        self.scalars: List[int] = []
        self.points: list = []
        self.g_scalars = None
        self.g0_scalar = self.w_scalar = self.u_scalar = 0

    def append_term(self, scalar: int, point):
        self.scalars.append(scalar % self.B.p)
        self.points.append(point)

    def add_constant_term(self, constant: int):              # adds `constant` to the coefficient of G_0
        self.g0_scalar = (self.g0_scalar + constant) % self.B.p

    def add_to_g_scalars(self, vec):
        self.g_scalars = vec if self.g_scalars is None else self.B.mul_add(self.g_scalars, 1, vec)

    def add_to_w_scalar(self, s: int): self.w_scalar = (self.w_scalar + s) % self.B.p
    def add_to_u_scalar(self, s: int): self.u_scalar = (self.u_scalar + s) % self.B.p

    def add_msm(self, other: "MSM"):
        self.scalars += other.scalars
        self.points += other.points
        if other.g_scalars is not None:
            self.add_to_g_scalars(other.g_scalars)
        self.add_constant_term(other.g0_scalar)
        self.add_to_w_scalar(other.w_scalar)
        self.add_to_u_scalar(other.u_scalar)

    def scale(self, factor: int):
        p = self.B.p
        self.scalars = [s * factor % p for s in self.scalars]
        if self.g_scalars is not None:
            self.g_scalars = self.B.mul_add(self.g_scalars, factor, self.B.vec([]))
        self.g0_scalar, self.w_scalar, self.u_scalar = (v * factor % p for v in (self.g0_scalar, self.w_scalar, self.u_scalar))

    def eval(self) -> bool:
        """True iff the whole combination is the identity"""
        g0, w, u = self.B.fixed_points()
        scalars = self.scalars + [self.g0_scalar, self.w_scalar, self.u_scalar]
        points = self.points + [g0, w, u]
        if self.g_scalars is not None:
            scalars.append(1)
            points.append(self.B.commit(self.g_scalars, 0))      # <g_scalars, G> over the resident generator table
        keep = [(s, pt) for s, pt in zip(scalars, points) if s and pt is not None]
        if not keep:
            return True
        return self.B.msm_points([s for s, _ in keep], [pt for _, pt in keep]) is None


class Guard:
    """commitment::Guard: the opening check with the n-term part (the challenges' s vector) still to be added"""

    def __init__(self, msm: MSM, neg_c: int, u: List[int]):
        self.msm, self.neg_c, self.u = msm, neg_c, u

    def use_challenges(self) -> MSM:
        self.msm.add_to_g_scalars(self.msm.B.ipa_s_vector(self.u, self.neg_c))
        return self.msm


def compute_b(x: int, u: List[int], p: int) -> int:
    """commitment::compute_b: prod over the rounds, last first, of (1 + u_j x^(2^i))"""
    tmp, cur = 1, x
    for u_j in reversed(u):
        tmp = tmp * (1 + u_j * cur) % p
        cur = cur * cur % p
    return tmp


def ipa_verify_proof(backend, msm: MSM, transcript: Blake2bRead, x: int, v: int) -> Guard:
    """poly::commitment::verify_proof: `msm` holds the commitment P that should open to v at x"""
    p = backend.p
    msm.add_constant_term(-v)
    s_commitment = transcript.read_point()
    xi = transcript.squeeze_challenge_scalar()
    msm.append_term(xi, s_commitment)
    z = transcript.squeeze_challenge_scalar()
    rounds = []
    for _ in range(backend.k):
        left, right = transcript.read_point(), transcript.read_point()
        rounds.append((left, right, transcript.squeeze_challenge_scalar()))
    u = []
    for left, right, u_j in rounds:
        if u_j == 0:
            raise VerifyError("Opening: a zero round challenge")
        msm.append_term(pow(u_j, -1, p), left)
        msm.append_term(u_j, right)
        u.append(u_j)
    c = transcript.read_scalar()
    neg_c = -c % p
    f = transcript.read_scalar()
    msm.add_to_u_scalar(neg_c * compute_b(x, u, p) % p * z)
    msm.add_to_w_scalar(-f)
    return Guard(msm, neg_c, u)


# ---- poly::multiopen::verify_proof -----------------------------------------------------------------------------------------------------
class VerifierQuery:
    """one (commitment, point, eval) triple; `commitment` is a point or an MSM (the folded h pieces), `key` identifies the
    commitment the way halo2's CommitmentReference compares them (by identity of the referenced object)"""

    def __init__(self, key, commitment, point: int, eval_: int):
        self.key, self.commitment, self.point, self.eval = key, commitment, point, eval_


def multiopen_verify_proof(backend, transcript: Blake2bRead, queries: List[VerifierQuery], msm: MSM) -> Guard:
    p = backend.p
    x_1 = transcript.squeeze_challenge_scalar()
    x_2 = transcript.squeeze_challenge_scalar()
    cmap, point_sets = construct_intermediate_sets(queries, lambda q: q.key, lambda q: q.point, lambda q: q.eval)
    q_commitments = [MSM(backend) for _ in point_sets]
    q_eval_sets = [[0] * len(ps) for ps in point_sets]
    for cd in cmap:
        s, cm = cd["set_index"], cd["commitment"].commitment
        q_commitments[s].scale(x_1)
        if isinstance(cm, MSM):
            q_commitments[s].add_msm(cm)
        else:
            q_commitments[s].append_term(1, cm)
        q_eval_sets[s] = [(a * x_1 + e) % p for a, e in zip(q_eval_sets[s], cd["evals"])]
    q_prime_commitment = transcript.read_point()
    x_3 = transcript.squeeze_challenge_scalar()
    u = [transcript.read_scalar() for _ in point_sets]
    x_4 = transcript.squeeze_challenge_scalar()
    msm_eval = 0
    for points, evals, proof_eval in zip(point_sets, q_eval_sets, u):
        r_poly = lagrange_interpolate(points, evals, p)
        r_eval = 0
        for coeff in reversed(r_poly):
            r_eval = (r_eval * x_3 + coeff) % p
        ev = (proof_eval - r_eval) % p
        for pt in points:
            d = (x_3 - pt) % p
            if d == 0:
                raise VerifyError("Opening: x_3 coincides with an opening point")
            ev = ev * pow(d, -1, p) % p
        msm_eval = (msm_eval * x_2 + ev) % p
    msm.append_term(1, q_prime_commitment)
    v = msm_eval
    for q_commitment, q_eval in zip(q_commitments, u):
        msm.scale(x_4)
        msm.add_msm(q_commitment)
        v = (v * x_4 + q_eval) % p
    return ipa_verify_proof(backend, msm, transcript, x_3, v)


# ---- plonk::VerificationStrategy ---------------------------------------------------------------------------------------------------------
class SingleVerifier:
    """plonk::SingleVerifier: evaluates the proof's MSM at once; verify_proof returns None or raises VerifyError"""

    def __init__(self, backend):
        self.msm = MSM(backend)

    def process(self, f: Callable[[MSM], Guard]):
        if not f(self.msm).use_challenges().eval():
            raise VerifyError("ConstraintSystemFailure: the opening proof does not verify")
        return None


class _Collect:
    """BatchVerifier's per-proof strategy: hands the finished MSM back instead of evaluating it"""

    def __init__(self, backend):
        self.msm = MSM(backend)

    def process(self, f):
        return f(self.msm).use_challenges()


class BatchVerifier:
    """plonk::BatchVerifier (test_utils.rs:56-61): proofs of ONE verifying key are checked together -- each proof's MSM is
    scaled by a random non-zero scalar and the sum is evaluated once.  finalize is False if any proof is malformed or the
    combined check fails (the reference then re-checks proof by proof with SingleVerifier, test_utils.rs:62-70)."""

    def __init__(self, rand: Optional[Callable[[], int]] = None):
        self.items = []
        self.rand = rand

    def add_proof(self, instances, proof: bytes):
        self.items.append((instances, bytes(proof)))

    def finalize(self, backend, vk) -> bool:
        p = backend.p
        rand = self.rand or (lambda: int.from_bytes(os.urandom(64), "little") % p)
        acc = MSM(backend)
        for instances, proof in self.items:
            try:
                msm = verify_proof(backend, vk, _Collect(backend), instances, Blake2bRead(proof, backend.q, p))
            except VerifyError:
                return False
            r = rand() % p
            while r == 0:
                r = rand() % p
            msm.scale(r)
            acc.add_msm(msm)
        return acc.eval()


# ---- plonk::verify_proof -------------------------------------------------------------------------------------------------------------------
def verify_proof(backend, vk, strategy, instances, transcript: Blake2bRead):
    """plonk::verify_proof for ONE circuit instance (what the reference passes, test_utils.rs:67, 111).  instances: one list of
    ints (or one backend vector) per instance column.  Returns strategy.process(...)'s value (None for SingleVerifier) or raises VerifyError."""
    B, cs = backend, vk.cs
    n, p, k = B.n, B.p, B.k
    if vk.k != k or B.j != vk.cs_degree:
        raise ValueError("the backend was built for another domain than the verifying key's")
    bf = cs.blinding_factors()
    usable = n - (bf + 1)
    chunk_len = vk.cs_degree - 2
    rot = B.rotate_omega
    if len(instances) != cs.num_instance:
        raise VerifyError("InvalidInstances")
    def column(col):            # lists of at most usable_rows ints, or backend vectors of n values (as plonk.create_proof takes them)
        if isinstance(col, (list, tuple)):
            if len(col) > usable:
                raise VerifyError("InstanceTooLarge")
            return B.vec(list(col))
        return B.vec(col)

    inst_values = [column(col) for col in instances]
    t = transcript
    t.common_scalar(vk.transcript_repr)
    inst_commitments = B.commit_lagrange_many(inst_values, [1] * len(inst_values))    # Blind::default() = 1
    for cm in inst_commitments:
        t.common_point(cm)
    adv_commitments = [t.read_point() for _ in range(cs.num_advice)]
    theta = t.squeeze_challenge_scalar()
    lk_permuted = [(t.read_point(), t.read_point()) for _ in cs.lookups]
    beta = t.squeeze_challenge_scalar()
    gamma = t.squeeze_challenge_scalar()
    n_sets = -(-len(cs.permutation) // chunk_len) if cs.permutation else 0
    perm_commitments = [t.read_point() for _ in range(n_sets)]
    lk_products = [t.read_point() for _ in cs.lookups]
    random_commitment = t.read_point()
    y = t.squeeze_challenge_scalar()
    h_pieces = [t.read_point() for _ in range(vk.cs_degree - 1)]
    x = t.squeeze_challenge_scalar()
    q_i, q_a, q_f = cs.queries[INSTANCE], cs.queries[ADVICE], cs.queries[FIXED]
    evals = {INSTANCE: [t.read_scalar() for _ in q_i], ADVICE: [t.read_scalar() for _ in q_a], FIXED: [t.read_scalar() for _ in q_f]}
    random_eval = t.read_scalar()
    sigma_evals = [t.read_scalar() for _ in cs.permutation]
    perm_evals = []                                            # (eval, next_eval, last_eval or None) per set
    for i in range(n_sets):
        e, e_next = t.read_scalar(), t.read_scalar()
        perm_evals.append((e, e_next, t.read_scalar() if i + 1 < n_sets else None))
    lk_evals = [tuple(t.read_scalar() for _ in range(5)) for _ in cs.lookups]      # product, product_next, a, a_inv, s

    # ---- the vanishing identity at x: h(x) (x^n - 1) = sum_i y^i expr_i(x) --------------------------------------------------
    xn = pow(x, n, p)
    if xn == 1:
        raise VerifyError("ConstraintSystemFailure: the evaluation point lies in the domain")
    n_inv, barycentric = pow(n, -1, p), (xn - 1) % p

    def l_at(rotation):                # the Lagrange basis polynomial of row `rotation mod n`, evaluated at x
        w = rot(1, rotation)
        return barycentric * n_inv % p * w % p * pow((x - w) % p, -1, p) % p

    l_0, l_last = l_at(0), l_at(-(bf + 1))
    l_blind = sum(l_at(-r) for r in range(1, bf + 1)) % p
    active = (1 - l_last - l_blind) % p

    def query(q):
        return evals[q.kind][cs.query_index(q.kind, q.column, q.rotation)]

    terms = [evaluate_expression(g, p, query) for g in cs.gates]
    if n_sets:
        terms.append(l_0 * (1 - perm_evals[0][0]) % p)
        z_last = perm_evals[-1][0]
        terms.append(l_last * (z_last * z_last - z_last) % p)
        for i in range(1, n_sets):
            terms.append(l_0 * (perm_evals[i][0] - perm_evals[i - 1][2]) % p)
        for i in range(n_sets):
            left, right = perm_evals[i][1], perm_evals[i][0]
            cur_delta = beta * x % p * pow(B.delta, i * chunk_len, p) % p
            for off, (kind, c) in enumerate(cs.permutation[i * chunk_len:(i + 1) * chunk_len]):
                value = evals[kind][cs.query_index(kind, c, 0)]
                left = left * ((value + beta * sigma_evals[i * chunk_len + off] + gamma) % p) % p
                right = right * ((value + cur_delta + gamma) % p) % p
                cur_delta = cur_delta * B.delta % p
            terms.append((left - right) * active % p)
    for (inputs, tables), (z, z_next, a, a_inv, s) in zip(cs.lookups, lk_evals):
        def compress(es):
            acc = 0
            for e in es:
                acc = (acc * theta + evaluate_expression(e, p, query)) % p
            return acc
        terms.append(l_0 * (1 - z) % p)
        terms.append(l_last * (z * z - z) % p)
        terms.append((z_next * (a + beta) % p * (s + gamma) - z * (compress(inputs) + beta) % p * (compress(tables) + gamma)) % p * active % p)
        terms.append(l_0 * (a - s) % p)
        terms.append((a - s) * (a - a_inv) % p * active % p)
    expected_h = 0
    for v in terms:
        expected_h = (expected_h * y + v) % p
    expected_h = expected_h * pow(xn - 1, -1, p) % p
    h_commitment = MSM(B)                                       # sum_i x^(n i) h_i, left as an MSM
    for piece in reversed(h_pieces):
        h_commitment.scale(xn)
        h_commitment.append_term(1, piece)

    # ---- the queries, in the prover's order (plonk/verifier.rs) ---------------------------------------------------------------
    x_next, x_inv, x_last = rot(x, 1), rot(x, -1), rot(x, -(bf + 1))
    Q: List[VerifierQuery] = []
    for (c, r), e in zip(q_i, evals[INSTANCE]):
        Q.append(VerifierQuery(("instance", c), inst_commitments[c], rot(x, r), e))
    for (c, r), e in zip(q_a, evals[ADVICE]):
        Q.append(VerifierQuery(("advice", c), adv_commitments[c], rot(x, r), e))
    for i in range(n_sets):
        Q.append(VerifierQuery(("perm_z", i), perm_commitments[i], x, perm_evals[i][0]))
        Q.append(VerifierQuery(("perm_z", i), perm_commitments[i], x_next, perm_evals[i][1]))
    for i in reversed(range(n_sets - 1)):
        Q.append(VerifierQuery(("perm_z", i), perm_commitments[i], x_last, perm_evals[i][2]))
    for li, ((cm_a, cm_s), cm_z, (z, z_next, a, a_inv, s)) in enumerate(zip(lk_permuted, lk_products, lk_evals)):
        Q.append(VerifierQuery(("lookup_z", li), cm_z, x, z))
        Q.append(VerifierQuery(("lookup_a", li), cm_a, x, a))
        Q.append(VerifierQuery(("lookup_s", li), cm_s, x, s))
        Q.append(VerifierQuery(("lookup_a", li), cm_a, x_inv, a_inv))
        Q.append(VerifierQuery(("lookup_z", li), cm_z, x_next, z_next))
    for (c, r), e in zip(q_f, evals[FIXED]):
        Q.append(VerifierQuery(("fixed", c), vk.fixed_commitments[c], rot(x, r), e))
    for i, e in enumerate(sigma_evals):
        Q.append(VerifierQuery(("sigma", i), vk.permutation_commitments[i], x, e))
    Q.append(VerifierQuery(("h",), h_commitment, x, expected_h))
    Q.append(VerifierQuery(("random",), random_commitment, x, random_eval))
    return strategy.process(lambda msm: multiopen_verify_proof(B, t, Q, msm))
