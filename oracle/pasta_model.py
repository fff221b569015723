"""Oracle tier (a): Python big-int golden model of the prover hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may use it, as the checker.

PARITY UNPINNED: the algorithms restated here live in third-party crates that
are absent from /root/reference (halo2_proofs 0.2.0 @ a95945254dcc, Cargo.lock:619-621;
pasta_curves 0.4.1, Cargo.lock:847-849) and the reference's tests hold no golden
prover outputs (src/test_utils.rs:56-70 only checks accept/reject).  What pins this
model is mathematics (every function has a unique answer) plus the pasta constants
re-derived in SURVEY.md Appendix A.  Call sites that select the instantiation:
src/test_utils.rs:2,12,21,40 (circuit field Fp, commitments on Vesta = EqAffine).

Restated (published algorithm, from upstream source as remembered):
  * pasta_curves::fields::{Fp,Fq}      -> class Field   (Montgomery R = 2^256, to_repr LE)
  * pasta_curves::curves::{Ep,Eq}      -> class Curve   (y^2 = x^3 + 5, compressed encoding)
  * halo2_proofs::arithmetic::best_fft -> best_fft      (bit-reverse + radix-2 DIT)
  * halo2_proofs::arithmetic::best_multiexp -> best_multiexp (result only; naive double-and-add
    and a Pippenger with the reference's window rule)
  * halo2_proofs::poly::EvaluationDomain -> class EvaluationDomain
"""
from __future__ import annotations

import math

MASK64 = (1 << 64) - 1


class Field:
    """Prime field of the Pasta cycle; elements are plain Python ints in [0, p)."""

    def __init__(self, name: str, p: int, zeta: int):
        self.name = name
        self.p = p
        self.R = (1 << 256) % p
        self.R2 = pow(1 << 256, 2, p)
        self.R3 = pow(1 << 256, 3, p)
        self.Rinv = pow(self.R, -1, p)
        self.INV64 = (-pow(p, -1, 1 << 64)) % (1 << 64)
        self.INV32 = (-pow(p, -1, 1 << 32)) % (1 << 32)
        self.S = 32
        self.T = (p - 1) >> 32
        self.GENERATOR = 5
        self.ROOT_OF_UNITY = pow(5, self.T, p)
        self.ZETA = zeta
        self.DELTA = pow(5, 1 << 32, p)
        self.TWO_INV = pow(2, -1, p)
        assert (p - 1) % (1 << 32) == 0 and self.T % 2 == 1
        assert pow(self.ROOT_OF_UNITY, 1 << 31, p) == p - 1
        assert pow(zeta, 3, p) == 1 and zeta != 1

    # -- element ops ------------------------------------------------------------------
    def add(self, a, b): return (a + b) % self.p
    def sub(self, a, b): return (a - b) % self.p
    def neg(self, a): return (-a) % self.p
    def mul(self, a, b): return (a * b) % self.p
    def sqr(self, a): return (a * a) % self.p
    def inv(self, a): return pow(a, self.p - 2, self.p)   # ff::Field::invert: 0 -> None; here 0 -> 0
    def pow(self, a, e): return pow(a, e, self.p)

    def sqrt(self, a):
        """Tonelli-Shanks; returns None for a non-residue."""
        p = self.p
        if a == 0:
            return 0
        if pow(a, (p - 1) // 2, p) != 1:
            return None
        # p - 1 = 2^S * T
        z = self.ROOT_OF_UNITY          # generator of the 2^S subgroup
        m = self.S
        c = z
        t = pow(a, self.T, p)
        r = pow(a, (self.T + 1) // 2, p)
        while t != 1:
            i, t2 = 0, t
            while t2 != 1:
                t2 = t2 * t2 % p
                i += 1
            b = pow(c, 1 << (m - i - 1), p)
            m = i
            c = b * b % p
            t = t * c % p
            r = r * b % p
        return r

    # -- representations --------------------------------------------------------------
    def to_mont(self, a): return (a * self.R) % self.p
    def from_mont(self, a): return (a * self.Rinv) % self.p

    @staticmethod
    def limbs(a):
        return [(a >> (64 * i)) & MASK64 for i in range(4)]

    @staticmethod
    def from_limbs(l):
        return sum(int(x) << (64 * i) for i, x in enumerate(l))

    def to_repr(self, a) -> bytes:
        """PrimeField::to_repr: 32-byte little-endian canonical."""
        return int(a).to_bytes(32, "little")

    def from_repr(self, b: bytes):
        v = int.from_bytes(b, "little")
        return v if v < self.p else None

    def root_of_unity(self, log_n: int):
        """omega with exact order 2^log_n: ROOT_OF_UNITY^(2^(S-log_n))."""
        assert 0 <= log_n <= self.S
        return pow(self.ROOT_OF_UNITY, 1 << (self.S - log_n), self.p)


P_MOD = 0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001
Q_MOD = 0x40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001

Fp = Field("Fp", P_MOD, 0x12ccca834acdba712caad5dc57aab1b01d1f8bd237ad31491dad5ebdfdfe4ab9)
Fq = Field("Fq", Q_MOD, 0x06819a58283e528e511db4d81cf70f5a0fed467d47c033af2aa9d2e050aa0e4f)


class Curve:
    """y^2 = x^3 + 5 over `base`, prime order = |scalar field|.  Affine points are (x, y) tuples,
    the identity is None."""

    B = 5

    def __init__(self, name, base: Field, scalar: Field):
        self.name, self.base, self.scalar = name, base, scalar
        self.G = (base.p - 1, 2)                       # pasta generator (-1, 2)
        assert self.on_curve(self.G)

    def on_curve(self, P):
        if P is None:
            return True
        x, y = P
        p = self.base.p
        return (y * y - x * x * x - self.B) % p == 0

    def neg(self, P):
        return None if P is None else (P[0], (-P[1]) % self.base.p)

    def add(self, P, Q):
        p = self.base.p
        if P is None: return Q
        if Q is None: return P
        x1, y1 = P
        x2, y2 = Q
        if x1 == x2:
            if (y1 + y2) % p == 0:
                return None
            lam = (3 * x1 * x1) * pow(2 * y1, -1, p) % p
        else:
            lam = (y2 - y1) * pow(x2 - x1, -1, p) % p
        x3 = (lam * lam - x1 - x2) % p
        y3 = (lam * (x1 - x3) - y1) % p
        return (x3, y3)

    def double(self, P): return self.add(P, P)

    # Jacobian ops (a = 0) for speed: (X, Y, Z), identity Z = 0
    def _jdouble(self, P):
        X, Y, Z = P
        p = self.base.p
        if Z == 0: return P
        A = X * X % p; Bq = Y * Y % p; C = Bq * Bq % p
        D = 2 * ((X + Bq) * (X + Bq) - A - C) % p
        E = 3 * A % p; F = E * E % p
        X3 = (F - 2 * D) % p
        Y3 = (E * (D - X3) - 8 * C) % p
        Z3 = 2 * Y * Z % p
        return (X3, Y3, Z3)

    def _jadd_affine(self, P, Q):
        if Q is None: return P
        X1, Y1, Z1 = P
        p = self.base.p
        if Z1 == 0: return (Q[0], Q[1], 1)
        x2, y2 = Q
        Z1Z1 = Z1 * Z1 % p
        U2 = x2 * Z1Z1 % p
        S2 = y2 * Z1 * Z1Z1 % p
        if U2 == X1:
            if S2 == Y1: return self._jdouble(P)
            return (0, 1, 0)
        H = (U2 - X1) % p; HH = H * H % p; HHH = H * HH % p
        r = (S2 - Y1) % p
        V = X1 * HH % p
        X3 = (r * r - HHH - 2 * V) % p
        Y3 = (r * (V - X3) - Y1 * HHH) % p
        Z3 = Z1 * H % p
        return (X3, Y3, Z3)

    def _to_affine(self, P):
        X, Y, Z = P
        p = self.base.p
        if Z == 0: return None
        zi = pow(Z, -1, p); zi2 = zi * zi % p
        return (X * zi2 % p, Y * zi2 * zi % p)

    def mul(self, k, P):
        k %= self.scalar.p
        if P is None or k == 0: return None
        acc = (0, 1, 0)
        for bit in bin(k)[2:]:
            acc = self._jdouble(acc)
            if bit == "1":
                acc = self._jadd_affine(acc, P)
        return self._to_affine(acc)

    def compress(self, P) -> bytes:
        """GroupEncoding::to_bytes (pasta_curves curves.rs): x LE, bit 255 = y & 1; identity = 32 zero bytes."""
        if P is None:
            return bytes(32)
        x, y = P
        b = bytearray(int(x).to_bytes(32, "little"))
        b[31] |= (y & 1) << 7
        return bytes(b)

    def decompress(self, b: bytes):
        if b == bytes(32): return None
        sign = b[31] >> 7
        x = int.from_bytes(b, "little") & ((1 << 255) - 1)
        y = self.base.sqrt((x * x * x + self.B) % self.base.p)
        assert y is not None
        if (y & 1) != sign: y = self.base.p - y
        return (x, y)

    def naive_msm(self, scalars, bases):
        acc = None
        for s, P in zip(scalars, bases):
            acc = self.add(acc, self.mul(s, P))
        return acc

    def best_multiexp(self, scalars, bases):
        """Result-equivalent restatement of halo2_proofs::arithmetic::best_multiexp (single thread):
        c = ceil(ln n) (1 if n<4, 3 if n<32), 256/c+1 unsigned segments, running-sum bucket reduction."""
        n = len(bases)
        assert len(scalars) == n
        c = 1 if n < 4 else 3 if n < 32 else math.ceil(math.log(n))
        segments = 256 // c + 1
        acc = (0, 1, 0)
        for seg in reversed(range(segments)):
            for _ in range(c):
                acc = self._jdouble(acc)
            buckets = [(0, 1, 0)] * ((1 << c) - 1)
            for s, P in zip(scalars, bases):
                d = (s >> (seg * c)) & ((1 << c) - 1)
                if d:
                    buckets[d - 1] = self._jadd_affine(buckets[d - 1], P)
            running = None
            for b in reversed(buckets):
                running = self.add(running, self._to_affine(b))
                acc = self._jadd_affine(acc, running)
        return self._to_affine(acc)


# Pallas: coordinates in Fp, scalars in Fq.  Vesta: coordinates in Fq, scalars in Fp.
Pallas = Curve("pallas", Fp, Fq)
Vesta = Curve("vesta", Fq, Fp)


# ---------------------------------------------------------------------------------------------
# halo2_proofs::arithmetic::best_fft  (field instance)
# ---------------------------------------------------------------------------------------------
def bitreverse(n, l):
    r = 0
    for _ in range(l):
        r = (r << 1) | (n & 1)
        n >>= 1
    return r


def best_fft(F: Field, a, omega, log_n):
    """In-place semantics of best_fft: natural in, natural out, A[k] = sum_j a[j] omega^(jk)."""
    n = 1 << log_n
    assert len(a) == n
    p = F.p
    a = list(a)
    for k in range(n):
        rk = bitreverse(k, log_n)
        if k < rk:
            a[k], a[rk] = a[rk], a[k]
    tw = [1] * max(n // 2, 1)
    for i in range(1, n // 2):
        tw[i] = tw[i - 1] * omega % p
    chunk, tchunk = 2, n // 2
    for _ in range(log_n):
        half = chunk // 2
        for s in range(0, n, chunk):
            for i in range(half):
                t = a[s + half + i] * tw[i * tchunk] % p
                u = a[s + i]
                a[s + i] = (u + t) % p
                a[s + half + i] = (u - t) % p
        chunk *= 2
        tchunk //= 2
    return a


def best_fft_group(C: "Curve", a, omega, log_n):
    """best_fft instantiated over curve points (halo2_proofs::arithmetic::best_fft is generic over `Group`; Params::new uses
    it on Vec<C::Curve>): the same butterflies, `t = a[hi] * twiddle` being a scalar multiplication.  Affine in / out."""
    n = 1 << log_n
    assert len(a) == n
    p = C.scalar.p
    a = list(a)
    for k in range(n):
        rk = bitreverse(k, log_n)
        if k < rk:
            a[k], a[rk] = a[rk], a[k]
    tw = [1] * max(n // 2, 1)
    for i in range(1, n // 2):
        tw[i] = tw[i - 1] * omega % p
    chunk, tchunk = 2, n // 2
    for _ in range(log_n):
        half = chunk // 2
        for s in range(0, n, chunk):
            for i in range(half):
                w = tw[i * tchunk]
                t = a[s + half + i] if w == 1 else C.mul(w, a[s + half + i])
                u = a[s + i]
                a[s + i] = C.add(u, t)
                a[s + half + i] = C.add(u, C.neg(t))
        chunk *= 2
        tchunk //= 2
    return a


def naive_dft(F: Field, a, omega):
    n = len(a)
    return [sum(a[j] * pow(omega, j * k, F.p) for j in range(n)) % F.p for k in range(n)]


# ---------------------------------------------------------------------------------------------
# halo2_proofs::poly::EvaluationDomain
# ---------------------------------------------------------------------------------------------
class EvaluationDomain:
    """Restates EvaluationDomain::new(j, k) and the coset transforms (poly/domain.rs, halo2 0.2.0)."""

    def __init__(self, F: Field, j: int, k: int):
        self.F, self.j, self.k = F, j, k
        p = F.p
        self.quotient_poly_degree = j - 1
        self.n = 1 << k
        ek = k
        while (1 << ek) < self.n * self.quotient_poly_degree:
            ek += 1
        self.extended_k = ek
        self.extended_omega = F.root_of_unity(ek)
        self.omega = pow(self.extended_omega, 1 << (ek - k), p)
        self.omega_inv = F.inv(self.omega)
        self.extended_omega_inv = F.inv(self.extended_omega)
        self.g_coset = F.ZETA
        self.g_coset_inv = F.sqr(F.ZETA)
        orig = pow(F.ZETA, self.n, p)
        step = pow(self.extended_omega, self.n, p)
        t, cur = [], orig
        while True:
            t.append(cur)
            cur = cur * step % p
            if cur == orig:
                break
        assert len(t) == 1 << (ek - k)
        self.t_evaluations = [F.inv((x - 1) % p) for x in t]    # stored inverted
        self.ifft_divisor = F.inv(self.n % p)
        self.extended_ifft_divisor = F.inv((1 << ek) % p)
        self.barycentric_weight = F.inv(self.n % p)

    def extended_len(self): return 1 << self.extended_k

    def lagrange_to_coeff(self, a):
        out = best_fft(self.F, a, self.omega_inv, self.k)
        return [x * self.ifft_divisor % self.F.p for x in out]

    def coeff_to_lagrange(self, a):
        return best_fft(self.F, a, self.omega, self.k)

    def _distribute_powers_zeta(self, a, into_coset):
        cp = [self.g_coset, self.g_coset_inv] if into_coset else [self.g_coset_inv, self.g_coset]
        p = self.F.p
        return [x if i % 3 == 0 else x * cp[i % 3 - 1] % p for i, x in enumerate(a)]

    def coeff_to_extended(self, a):
        assert len(a) == self.n
        a = self._distribute_powers_zeta(a, True)
        a = a + [0] * (self.extended_len() - self.n)
        return best_fft(self.F, a, self.extended_omega, self.extended_k)

    def extended_to_coeff(self, a):
        assert len(a) == self.extended_len()
        a = best_fft(self.F, a, self.extended_omega_inv, self.extended_k)
        a = [x * self.extended_ifft_divisor % self.F.p for x in a]
        a = self._distribute_powers_zeta(a, False)
        return a[: self.n * self.quotient_poly_degree]

    def divide_by_vanishing_poly(self, a):
        assert len(a) == self.extended_len()
        m = len(self.t_evaluations)
        return [x * self.t_evaluations[i % m] % self.F.p for i, x in enumerate(a)]

    def rotate_extended(self, a, rotation):
        """rotate_extended: new[i] = old[i + rotation * 2^(ext_k-k)] (cyclic)."""
        n = len(a)
        s = (rotation * (1 << (self.extended_k - self.k))) % n
        return a[s:] + a[:s]

    def rotate_omega(self, v, rotation):
        return v * pow(self.omega if rotation >= 0 else self.omega_inv, abs(rotation), self.F.p) % self.F.p


def eval_polynomial(F: Field, coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % F.p
    return acc


# ---------------------------------------------------------------------------------------------
# halo2_proofs::poly::Evaluator::evaluate  (poly/evaluator.rs, halo2 0.2.0) -- restated from memory, UNVERIFIED:
# the dependency's source is not under /root/reference.  Walks an Ast (any objects whose CLASS NAMES are the
# halo2 node names Poly / Add / Mul / Scale / DistributePowers / LinearTerm / ConstantTerm) and returns the
# 2^extended_k values, one per point zeta * extended_omega^i of the extended coset.
#   Poly(index, rotation)      -> polys[index] rotated by rotation * 2^(extended_k - k)   (rotate_extended)
#   Add / Mul                  -> pointwise
#   Scale(a, s)                -> s * a
#   DistributePowers(ts, base) -> fold(0, |acc, t| acc * base + t)
#   LinearTerm(s)              -> s * zeta * extended_omega^i
#   ConstantTerm(s)            -> s
# ---------------------------------------------------------------------------------------------
def evaluate_ast(dom: EvaluationDomain, ast, polys):
    p = dom.F.p
    rows = dom.extended_len()
    kind = type(ast).__name__
    if kind == "Poly":
        return dom.rotate_extended(list(polys[ast.index]), ast.rotation)
    if kind == "Add":
        a, b = evaluate_ast(dom, ast.a, polys), evaluate_ast(dom, ast.b, polys)
        return [(x + y) % p for x, y in zip(a, b)]
    if kind == "Mul":
        a, b = evaluate_ast(dom, ast.a, polys), evaluate_ast(dom, ast.b, polys)
        return [x * y % p for x, y in zip(a, b)]
    if kind == "Scale":
        s = ast.scalar % p
        return [x * s % p for x in evaluate_ast(dom, ast.a, polys)]
    if kind == "DistributePowers":
        base = evaluate_ast(dom, ast.base, polys)
        acc = [0] * rows
        for t in ast.terms:
            tv = evaluate_ast(dom, t, polys)
            acc = [(x * b + y) % p for x, b, y in zip(acc, base, tv)]
        return acc
    if kind == "LinearTerm":
        out, cur = [], dom.g_coset
        for _ in range(rows):
            out.append(cur * ast.scalar % p)
            cur = cur * dom.extended_omega % p
        return out
    if kind == "ConstantTerm":
        return [ast.scalar % p] * rows
    raise TypeError(f"unknown Ast node {kind}")


def run_program(dom: EvaluationDomain, code, consts, polys, coset=-1):
    """Reference interpreter of the straight-line program format of include/tr_prover.h (trp_dev_quotient_eval),
    used by the CPU tests to check the host-side Ast compiler without a GPU.  coset = -1: whole extended domain;
    coset = j: rows of the j-th size-n coset, polys given on that coset, result returned for that coset only."""
    p = dom.F.p
    period = 1 << (dom.extended_k - dom.k)
    rows = dom.extended_len() if coset < 0 else dom.n
    step = period if coset < 0 else 1
    out = [None] * rows
    for row in range(rows):
        g = row if coset < 0 else row * period + coset
        regs = {}
        for op, dst, a, b in code:
            op, dst, a, b = int(op), int(dst), int(a), int(b)
            if op == 0:
                rot = b - (1 << 32) if b >= (1 << 31) else b
                regs[dst] = polys[a][(row + rot * step) % rows]
            elif op == 1: regs[dst] = consts[a]
            elif op == 2: regs[dst] = (regs[a] + regs[b]) % p
            elif op == 3: regs[dst] = (regs[a] - regs[b]) % p
            elif op == 4: regs[dst] = regs[a] * regs[b] % p
            elif op == 5: regs[dst] = -regs[a] % p
            elif op == 6: regs[dst] = regs[a] * regs[a] % p
            elif op == 7: regs[dst] = 2 * regs[a] % p
            elif op == 8: regs[dst] = dom.g_coset * pow(dom.extended_omega, g, p) % p
            elif op == 9: out[row] = regs[a]
            elif op == 10: regs[dst] = regs[a] * consts[b] % p
            elif op == 11: regs[dst] = (regs[a] + consts[b]) % p
            elif op == 12: regs[dst] = (regs[a] - consts[b]) % p
            else: raise ValueError(f"bad opcode {op}")
    return out


# ---- the field work between the hot kernels (halo2_proofs 0.2.0, restated from the published
# source as remembered; the reference reaches it through create_proof, src/test_utils.rs:41,96) ------------------------------
def batch_invert(F: Field, vals):
    """ff::BatchInvert: Montgomery's trick over the NON-ZERO entries; zeros are left untouched."""
    acc, prefix = 1, []
    for v in vals:
        prefix.append(acc)
        if v != 0:
            acc = acc * v % F.p
    inv = F.inv(acc)
    out = list(vals)
    for i in range(len(vals) - 1, -1, -1):
        if vals[i] == 0:
            continue
        out[i] = inv * prefix[i] % F.p
        inv = inv * vals[i] % F.p
    return out


def permutation_commit(F: Field, omega, n, values, permutations, beta, gamma, chunk_len, blinding_factors, rand, after_chunk=None):
    """plonk::permutation::prover::Argument::commit: one grand-product column Z per chunk of `chunk_len` columns.
    values[c] / permutations[c]: the n Lagrange values of column c and of its sigma polynomial; rand() draws a scalar
    (blinding rows of every Z, in order).  Returns the Z columns (blind scalars / commitments are the caller's)."""
    p = F.p
    deltaomega, last_z, sets = 1, 1, []
    for lo in range(0, len(values), chunk_len):
        cols, perms = values[lo:lo + chunk_len], permutations[lo:lo + chunk_len]
        modified = [1] * n
        for col, perm in zip(cols, perms):
            for i in range(n):
                modified[i] = modified[i] * ((beta * perm[i] + gamma + col[i]) % p) % p
        modified = batch_invert(F, modified)
        for col in cols:
            dw = deltaomega
            for i in range(n):
                modified[i] = modified[i] * ((dw * beta + gamma + col[i]) % p) % p
                dw = dw * omega % p
            deltaomega = deltaomega * F.DELTA % p
        z = [last_z]
        for row in range(1, n):
            z.append(z[row - 1] * modified[row - 1] % p)
        for i in range(n - blinding_factors, n):
            z[i] = rand()
        last_z = z[n - (blinding_factors + 1)]
        sets.append(z)
        if after_chunk is not None:          # create_proof draws the chunk's blind and commits before the next chunk starts
            after_chunk(z)
    return sets


def permute_expression_pair(F: Field, input_expression, table_expression, usable_rows):
    """plonk::lookup::prover::permute_expression_pair without the trailing blinding rows: returns
    (permuted_input, permuted_table) over the usable rows, or None where halo2 returns ConstraintSystemFailure."""
    permuted_input = sorted(input_expression[:usable_rows])          # Ord for Fp/Fq = canonical integer order
    leftover = {}
    for v in table_expression[:usable_rows]:
        leftover[v] = leftover.get(v, 0) + 1
    permuted_table = [0] * usable_rows
    repeated_rows = []
    for row, v in enumerate(permuted_input):
        if row == 0 or v != permuted_input[row - 1]:
            permuted_table[row] = v
            if leftover.get(v, 0) > 0:
                leftover[v] -= 1
            else:
                return None
        else:
            repeated_rows.append(row)
    for coeff in sorted(leftover):                                   # BTreeMap iteration order
        for _ in range(leftover[coeff]):
            permuted_table[repeated_rows.pop()] = coeff
    assert not repeated_rows
    return permuted_input, permuted_table


def lookup_commit_product(F: Field, n, compressed_input, compressed_table, permuted_input, permuted_table, beta, gamma,
                          blinding_factors, rand):
    """plonk::lookup::prover::Permuted::commit_product: the lookup grand product Z (n values, blinding tail from rand())."""
    p = F.p
    prod = [(beta + permuted_input[i]) * (gamma + permuted_table[i]) % p for i in range(n)]
    prod = batch_invert(F, prod)
    for i in range(n):
        prod[i] = prod[i] * ((compressed_input[i] + beta) % p) % p * ((compressed_table[i] + gamma) % p) % p
    z, state = [], 1
    for cur in [1] + prod:
        state = state * cur % p
        z.append(state)
    z = z[:n - blinding_factors] + [rand() for _ in range(blinding_factors)]
    return z


# ---- the opening phase (halo2_proofs 0.2.0 arithmetic.rs / poly/commitment/prover.rs) -------------
def compute_inner_product(F: Field, a, b):
    assert len(a) == len(b)
    acc = 0
    for x, y in zip(a, b):
        acc = (acc + x * y) % F.p
    return acc


def kate_division(F: Field, a, b):
    """arithmetic::kate_division: divide a(X) by (X - b), discarding the remainder; len(a) - 1 coefficients."""
    b = (-b) % F.p
    q = [0] * (len(a) - 1)
    tmp = 0
    for i in range(len(a) - 1, 0, -1):
        lead = (a[i] - tmp) % F.p
        q[i - 1] = lead
        tmp = lead * b % F.p
    return q


def parallel_generator_collapse(C: "Curve", g, challenge):
    half = len(g) // 2
    return [C.add(g[i], C.mul(challenge, g[i + half])) for i in range(half)]


def ipa_create_proof(C: "Curve", k, g, w, u, rand, transcript, p_poly, p_blind, x_3):
    """poly::commitment::prover::create_proof.  g: n affine points (None = identity), w, u: points; p_poly: n canonical
    coefficients.  transcript: write_point(P), write_scalar(s), squeeze_challenge_scalar() -> int; rand() -> scalar."""
    F = C.scalar
    p, n = F.p, 1 << k
    assert len(p_poly) == n
    s_poly = [rand() for _ in range(n)]
    s_at_x3 = eval_polynomial(F, s_poly, x_3)
    s_poly[0] = (s_poly[0] - s_at_x3) % p
    s_poly_blind = rand()
    transcript.write_point(C.best_multiexp(s_poly + [s_poly_blind], g + [w]))
    xi = transcript.squeeze_challenge_scalar()
    z = transcript.squeeze_challenge_scalar()
    p_prime = [(s * xi + c) % p for s, c in zip(s_poly, p_poly)]
    v = eval_polynomial(F, p_prime, x_3)
    p_prime[0] = (p_prime[0] - v) % p
    f = (s_poly_blind * xi + p_blind) % p
    b, cur = [], 1
    for _ in range(n):
        b.append(cur)
        cur = cur * x_3 % p
    g_prime = list(g)
    for j in range(k):
        half = 1 << (k - j - 1)
        l_j = C.best_multiexp(p_prime[half:], g_prime[:half])
        r_j = C.best_multiexp(p_prime[:half], g_prime[half:])
        value_l_j = compute_inner_product(F, p_prime[half:], b[:half])
        value_r_j = compute_inner_product(F, p_prime[:half], b[half:])
        l_rand, r_rand = rand(), rand()
        l_j = C.add(l_j, C.best_multiexp([value_l_j * z % p, l_rand], [u, w]))
        r_j = C.add(r_j, C.best_multiexp([value_r_j * z % p, r_rand], [u, w]))
        transcript.write_point(l_j)
        transcript.write_point(r_j)
        u_j = transcript.squeeze_challenge_scalar()
        u_j_inv = F.inv(u_j)
        for i in range(half):
            p_prime[i] = (p_prime[i] + p_prime[i + half] * u_j_inv) % p
            b[i] = (b[i] + b[i + half] * u_j) % p
        p_prime, b = p_prime[:half], b[:half]
        g_prime = parallel_generator_collapse(C, g_prime, u_j)
        f = (f + l_rand * u_j_inv + r_rand * u_j) % p
    assert len(p_prime) == 1
    transcript.write_scalar(p_prime[0])
    transcript.write_scalar(f)
