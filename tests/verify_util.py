"""Test infrastructure: run oracle/plonk_model.verify_proof (the independent restatement of halo2's verifier) on a proof made by
plonk.GpuBackend, with the verifier's size-n MSMs on the C++ oracle so that k = 20 verifies in seconds."""
import numpy as np

import oracle as O
import pasta_model as pm
import plonk_model as VM


class _Marker(list):
    """a list of points that remembers the limb array it was made from (the big MSMs skip the int -> limb conversion)"""
    limbs = None

    def __add__(self, other):
        out = _Marker(list.__add__(self, other)); out.limbs = self.limbs; return out


def oracle_params(be):
    """the GPU backend's Params (tested against params_model in test_gpu_params.py) in the oracle's representation"""
    pts = lambda arr: [None if not r.any() else tuple(be._ints(r.reshape(2, 4), be.q, be.Rqinv)) for r in np.asarray(arr).reshape(-1, 8)]
    g_l = np.concatenate([np.asarray(be.params.g_lagrange_points).reshape(-1, 8), np.asarray(be.params.w).reshape(1, 8)])
    g_c = np.asarray(be.params.g_points).reshape(-1, 8)
    gl = _Marker(pts(g_l[:-1])); gl.limbs = g_l
    gc = _Marker(pts(g_c)); gc.limbs = g_c
    return {"k": be.k, "n": be.n, "g": gc, "g_lagrange": gl, "w": pts(be.params.w)[0], "u": pts(be.params.u)[0]}


def fast_vesta():
    p = pm.Vesta.scalar.p
    bf, sf = O.BASE_FIELD[O.VESTA], O.SCALAR_FIELD[O.VESTA]

    class FastCurve(pm.Curve):
        def best_multiexp(self, scalars, bases):
            limbs = getattr(bases, "limbs", None)
            if limbs is None or len(bases) < 256:
                return super().best_multiexp(scalars, bases)
            nz = [i for i, s in enumerate(scalars) if s % p]
            if not nz:
                return None
            sc = O.to_mont(sf, O.ints_to_limbs([scalars[i] % p for i in nz]))
            out = O.msm(O.VESTA, sc, np.ascontiguousarray(limbs[nz]))
            if not out.any():
                return None
            x, y = O.limbs_to_ints(O.from_mont(bf, out.reshape(2, 4)))
            return (x, y)

    return FastCurve("vesta", pm.Fq, pm.Fp)


def verify(be, vk, instances, proof, curve=None):
    curve = curve or fast_vesta()
    ok = VM.verify_proof(curve, oracle_params(be), vk, instances, proof)
    return ok, (None if ok else VM.verify_proof.last_error)
