"""Shared helpers for the parity tests (oracle-side input generation)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle as O  # noqa: E402
import pasta_model as pm  # noqa: E402

CURVES = {O.PALLAS: pm.Pallas, O.VESTA: pm.Vesta}
FIELDS = {O.FP: pm.Fp, O.FQ: pm.Fq}


def generator(curve):
    bf = O.BASE_FIELD[curve]
    g = np.zeros(8, dtype=np.uint64)
    g[:4] = O.to_mont(bf, O.ints_to_limbs([O.MODULUS[bf] - 1]))[0]
    g[4:] = O.to_mont(bf, O.ints_to_limbs([2]))[0]
    return g


def make_points(curve, n, seed=21):
    """n pseudo-random affine points P0 + i*D (arithmetic progression of random multiples of G)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    q = O.MODULUS[O.SCALAR_FIELD[curve]]
    k0 = int.from_bytes(rng.bytes(32), "little") % q
    k1 = int.from_bytes(rng.bytes(32), "little") % q
    g = generator(curve)
    p0 = O.point_mul(curve, O.ints_to_limbs([k0]), g)
    d = O.point_mul(curve, O.ints_to_limbs([k1]), g)
    return O.points_progression(curve, p0, d, n)


def scalars_uniform(curve, n, seed=20):
    return O.random_field_mont(O.SCALAR_FIELD[curve], n, seed)


def scalars_tinyram(curve, n, seed=40):
    """TinyRAM-shaped column: 90 % {0,1}, 8 % < 2^32, 2 % uniform (SURVEY 8a MSM note); Montgomery form."""
    rng = np.random.Generator(np.random.PCG64(seed))
    f = O.SCALAR_FIELD[curve]
    canon = np.zeros((n, 4), dtype=np.uint64)
    kind = rng.random(n)
    canon[:, 0] = np.where(kind < 0.9, rng.integers(0, 2, n, dtype=np.uint64), rng.integers(0, 1 << 32, n, dtype=np.uint64))
    out = O.to_mont(f, canon)
    uni = O.random_field_mont(f, n, seed + 1)
    sel = kind >= 0.98
    out[sel] = uni[sel]
    return out


def affine_of(curve, jac):
    return O.jacobian_to_affine(curve, jac)
