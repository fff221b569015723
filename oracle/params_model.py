"""Oracle for SURVEY.md 8(f) row f3: halo2_proofs::poly::commitment::Params::new(k) and the pieces of pasta_curves it calls.

TEST INFRASTRUCTURE ONLY (see pasta_model.py's header): the product never imports this.

Reached from the reference at src/test_utils.rs:21,89 (`Params::<EqAffine>::new(k)`).  The code it runs lives in crates that
are not under /root/reference (halo2_proofs 0.2.0 @ a95945254dcc, Cargo.lock:619-621; pasta_curves 0.4.1, Cargo.lock:847-849);
restated from the published sources:

  * pasta_curves::hashtocurve::hash_to_field          -> hash_to_field   (expand_message_xmd over BLAKE2b-512, two
                                                          64-byte chunks, big-endian -> from_bytes_wide)
  * pasta_curves::hashtocurve::map_to_curve_simple_swu -> map_to_curve_simple_swu (draft-irtf-cfrg-hash-to-curve-10, 6.6.2)
  * pasta_curves::hashtocurve::iso_map                 -> iso_map (3-isogeny iso-Pallas -> Pallas, iso-Vesta -> Vesta)
  * CurveExt::hash_to_curve                            -> hash_to_curve
  * Params::new                                        -> params_new (g, g_lagrange by a group iFFT, w, u)

What pins this part (unlike the rest of the path, where the reference holds no vectors):
  1. the iso-curve coefficients a (recalled) give curves of the right ORDER (|iso-Pallas(Fp)| = q, |iso-Vesta(Fq)| = p);
  2. the 13 isogeny constants are DERIVED here by Velu's formulas from the unique rational 3-torsion subgroup and the
     isomorphism (x, y) -> (x / 9, y / 27) onto y^2 = x^3 + 5, and the Pallas set equals pasta_curves' published
     constants limb for limb (tests/test_params_cpu.py);
  3. pasta_curves' own known-answer test (hash_to_curve("z.cash:test")(b"Trans rights now!") on Pallas, Jacobian
     coordinates in its test module) is reproduced by hash_to_curve below (same test file).
"""
from __future__ import annotations

import hashlib

from pasta_model import Fp, Fq, Field, Curve, Pallas, Vesta, best_fft_group  # noqa: F401

ISO_A = {"pallas": 0x18354a2eb0ea8c9c49be2d7258370742b74134581a27a59f92bb4b0b657a014b,
         "vesta": 0x267f9b2ee592271a81639c4d96f787739673928c7d01b212c515ad7242eaa6b1}
ISO_B = 1265
SWU_Z = -13          # pasta_curves: Ep::Z = Eq::Z = -13


# ---- polynomial helpers over F_m (low -> high coefficient lists), only used to derive the isogeny ------------------------------
def _polymulmod(A, B, M, m):
    d = len(M) - 1
    R = [0] * (len(A) + len(B) - 1)
    for i, a in enumerate(A):
        if a:
            for j, b in enumerate(B):
                R[i + j] = (R[i + j] + a * b) % m
    for i in range(len(R) - 1, d - 1, -1):
        c = R[i]
        if c:
            for j in range(d + 1):
                R[i - d + j] = (R[i - d + j] - c * M[j]) % m
    R = R[:d]
    return R + [0] * (d - len(R))


def _polypow(B, e, M, m):
    d = len(M) - 1
    R = [1] + [0] * (d - 1)
    B = (B + [0] * d)[:d]
    while e:
        if e & 1:
            R = _polymulmod(R, B, M, m)
        B = _polymulmod(B, B, M, m)
        e >>= 1
    return R


def _trim(P):
    while P and P[-1] == 0:
        P = P[:-1]
    return P


def _polygcd(A, B, m):
    A, B = _trim(list(A)), _trim(list(B))
    while B:
        inv = pow(B[-1], -1, m)
        while len(A) >= len(B):
            c = A[-1] * inv % m
            s = len(A) - len(B)
            for j in range(len(B)):
                A[s + j] = (A[s + j] - c * B[j]) % m
            A = _trim(A)
            if not A:
                break
        A, B = B, A
    inv = pow(A[-1], -1, m)
    return [c * inv % m for c in A]


def _rational_roots(P, m):
    """roots in F_m of the monic polynomial P (Cantor-Zassenhaus, deterministic seed)"""
    import random
    rnd = random.Random(5)
    xp = _polypow([0, 1], m, P, m)
    xp[1] = (xp[1] - 1) % m
    g = _polygcd(P, xp, m)
    out = []

    def split(g):
        d = len(g) - 1
        if d == 0:
            return
        if d == 1:
            out.append((-g[0]) % m)
            return
        while True:
            R = _polypow([rnd.randrange(m), 1], (m - 1) // 2, g, m)
            R[0] = (R[0] - 1) % m
            h = _polygcd(g, R, m)
            if 0 < len(h) - 1 < d:
                A, Q = list(g), [0] * (d - (len(h) - 1) + 1)
                while len(A) >= len(h):
                    c, s = A[-1], len(A) - len(h)
                    Q[s] = c
                    for j in range(len(h)):
                        A[s + j] = (A[s + j] - c * h[j]) % m
                    A = A[:-1]
                split(h)
                split(Q)
                return

    split(g)
    return sorted(out)


def derive_isogeny_constants(F: Field, a: int, b: int = ISO_B):
    """The 13 constants of pasta_curves' iso_map, derived: E': y^2 = x^3 + a x + b has exactly one rational subgroup of order 3
    (x_Q = the rational root of the 3-division polynomial 3x^4 + 6a x^2 + 12b x - a^2); Velu's isogeny with that kernel lands on
    y^2 = x^3 + 5 * 3^6, and (x, y) -> (x / 3^2, y / 3^3) carries that to y^2 = x^3 + 5.
       x' = (c0 x^3 + c1 x^2 + c2 x + c3) / (x^2 + c4 x + c5)
       y' = y (c6 x^3 + c7 x^2 + c8 x + c9) / (x^3 + c10 x^2 + c11 x + c12)"""
    m = F.p
    i3 = pow(3, -1, m)
    psi3 = [(-a * a) * i3 % m, 12 * b * i3 % m, 6 * a * i3 % m, 0, 1]
    roots = _rational_roots(psi3, m)
    assert len(roots) == 1, "expected a unique rational 3-torsion subgroup"
    xq = roots[0]
    v = 2 * (3 * xq * xq + a) % m
    u = 4 * (xq ** 3 + a * xq + b) % m
    w = (u + xq * v) % m
    assert (a - 5 * v) % m == 0 and (b - 7 * w) % m == 5 * 729, "codomain is not y^2 = x^3 + 5 * 3^6"
    i9, i27 = pow(9, -1, m), pow(27, -1, m)
    c = [i9, -2 * xq * i9, (xq * xq + v) * i9, (u - v * xq) * i9,
         -2 * xq, xq * xq,
         i27, -3 * xq * i27, (3 * xq * xq - v) * i27, (-xq ** 3 + v * xq - 2 * u) * i27,
         -3 * xq, 3 * xq * xq, -xq ** 3]
    return [t % m for t in c]


_ISO_CACHE = {}


def isogeny_constants(curve: Curve):
    if curve.name not in _ISO_CACHE:
        _ISO_CACHE[curve.name] = derive_isogeny_constants(curve.base, ISO_A[curve.name])
    return _ISO_CACHE[curve.name]


# ---- pasta_curves::hashtocurve ---------------------------------------------------------------------------------------------------
def hash_to_field(F: Field, curve_id: str, domain_prefix: str, message: bytes):
    """Two field elements: expand_message_xmd (BLAKE2b, 64-byte output, empty personalisation, 128 zero bytes of Z_pad), DST =
    domain_prefix || "-" || curve_id || "_XMD:BLAKE2b_SSWU_RO_"; each 64-byte chunk is read BIG-endian and reduced."""
    assert len(domain_prefix) < 256 and 22 + len(curve_id) + len(domain_prefix) < 256
    dst = domain_prefix.encode() + b"-" + curve_id.encode() + b"_XMD:BLAKE2b_SSWU_RO_"
    dst_prime = dst + bytes([22 + len(curve_id) + len(domain_prefix)])
    H = lambda data: hashlib.blake2b(data, digest_size=64, person=bytes(16)).digest()
    b0 = H(bytes(128) + message + bytes([0, 128, 0]) + dst_prime)
    b1 = H(b0 + bytes([1]) + dst_prime)
    b2 = H(bytes(x ^ y for x, y in zip(b0, b1)) + bytes([2]) + dst_prime)
    return [int.from_bytes(b, "big") % F.p for b in (b1, b2)]


def map_to_curve_simple_swu(F: Field, a: int, b: int, z: int, u: int):
    """Simplified SWU onto y^2 = x^3 + a x + b (a, b != 0).  Affine (x, y); y's parity equals u's (sgn0)."""
    p = F.p
    z %= p
    z_u2 = z * u * u % p
    ta = (z_u2 * z_u2 + z_u2) % p
    num_x1 = b * (ta + 1) % p
    div = a * (z if ta == 0 else -ta) % p
    x1 = num_x1 * F.inv(div) % p
    gx1 = (x1 ** 3 + a * x1 + b) % p
    y1 = F.sqrt(gx1)
    if y1 is not None:
        x, y = x1, y1
    else:
        x = z_u2 * x1 % p
        y = F.sqrt((x ** 3 + a * x + b) % p)
        assert y is not None
    if (u & 1) != (y & 1):
        y = p - y if y else 0
    return (x, y)


def _iso_add(F: Field, a: int, P, Q):
    """affine addition on y^2 = x^3 + a x + b"""
    p = F.p
    if P is None:
        return Q
    if Q is None:
        return P
    (x1, y1), (x2, y2) = P, Q
    if x1 == x2:
        if (y1 + y2) % p == 0:
            return None
        lam = (3 * x1 * x1 + a) * F.inv(2 * y1 % p) % p
    else:
        lam = (y2 - y1) * F.inv((x2 - x1) % p) % p
    x3 = (lam * lam - x1 - x2) % p
    return (x3, (lam * (x1 - x3) - y1) % p)


def iso_map(F: Field, c, P):
    if P is None:
        return None
    p = F.p
    x, y = P
    num_x = ((c[0] * x + c[1]) * x + c[2]) * x + c[3]
    div_x = (x + c[4]) * x + c[5]
    num_y = (((c[6] * x + c[7]) * x + c[8]) * x + c[9]) * y
    div_y = ((x + c[10]) * x + c[11]) * x + c[12]
    if div_x % p == 0 or div_y % p == 0:      # the kernel of the isogeny
        return None
    return (num_x * F.inv(div_x % p) % p, num_y * F.inv(div_y % p) % p)


def hash_to_curve(curve: Curve, domain_prefix: str):
    """CurveExt::hash_to_curve(domain_prefix) -> closure message -> affine point (None = identity)."""
    F = curve.base
    a = ISO_A[curve.name]
    c = isogeny_constants(curve)

    def hasher(message: bytes):
        u0, u1 = hash_to_field(F, curve.name, domain_prefix, message)
        q0 = map_to_curve_simple_swu(F, a, ISO_B, SWU_Z, u0)
        q1 = map_to_curve_simple_swu(F, a, ISO_B, SWU_Z, u1)
        return iso_map(F, c, _iso_add(F, a, q0, q1))

    return hasher


# ---- halo2_proofs::poly::commitment::Params::new ---------------------------------------------------------------------------------
def params_generators(curve: Curve, n: int, start: int = 0):
    """g[i] = hash_to_curve("Halo2-Parameters")([0] ++ u32_le(i)), i in [start, start + n)"""
    h = hash_to_curve(curve, "Halo2-Parameters")
    return [h(bytes([0]) + (i & 0xFFFFFFFF).to_bytes(4, "little")) for i in range(start, start + n)]


def g_to_lagrange(curve: Curve, g, k: int):
    """The group iFFT of Params::new: best_fft over curve points with alpha^-1 = ROOT_OF_UNITY_INV^(2^(S-k)), then every point
    times TWO_INV^k.  g_lagrange[i] = sum_j L_i-coefficient_j * g[j], i.e. commit_lagrange(v) = commit(lagrange_to_coeff(v))."""
    Fs = curve.scalar
    alpha_inv = Fs.inv(Fs.root_of_unity(k))
    out = best_fft_group(curve, list(g), alpha_inv, k)
    minv = pow(Fs.TWO_INV, k, Fs.p)
    return [curve.mul(minv, P) for P in out]


def params_new(curve: Curve, k: int):
    """Params::new(k) -> dict(k, n, g, g_lagrange, w, u) with affine points"""
    n = 1 << k
    g = params_generators(curve, n)
    h = hash_to_curve(curve, "Halo2-Parameters")
    return {"k": k, "n": n, "g": g, "g_lagrange": g_to_lagrange(curve, g, k), "w": h(bytes([1])), "u": h(bytes([2]))}
