"""A REAL plonk.create_proof of a satisfiable circuit with the TinyRamCircuit's shape (tinyram_circuit.py: 263 advice / 94
instance / 23 fixed columns, 139 gates of degree <= 6, 31 lookups incl. the 95-wide dynamic one, 188 equality columns) on one
B200, checked by the oracle's independent verify_proof (its size-n MSMs run on the C++ oracle).
usage: python tests/gpu_tinyram_proof.py [k] [scale] [--no-verify]"""
import json
import os
import random
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE); sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import __graft_entry__ as ge
import oracle as O
import pasta_model as pm
import plonk_model as VM

args = [a for a in sys.argv[1:] if not a.startswith("--")]
k = int(args[0]) if args else 12
scale = float(args[1]) if len(args) > 1 else 1.0
verify = "--no-verify" not in sys.argv
import torch
pkg = ge.load_package()
PL = pkg.plonk
from tiny_ram_halo2_b200 import tinyram_circuit
C = pm.Vesta
p = C.scalar.p
ctx = pkg.Context(0, pkg.VESTA)


class Rng:
    """the caller's RNG: scalar draws from Python's Mersenne twister, bulk draws (random polynomials) from PCG64"""
    def __init__(self, seed):
        self.r, self.g = random.Random(seed), np.random.Generator(np.random.PCG64(seed))
    def __call__(self):
        return self.r.randrange(p)
    def vector(self, n):
        a = self.g.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
        a[:, 3] &= np.uint64((1 << 62) - 1)            # below 2^254: valid Montgomery representations of uniform elements
        return a


t0 = time.perf_counter()
be = PL.GpuBackend(ctx, k, 6)
torch.cuda.synchronize()
t_params = time.perf_counter() - t0
t0 = time.perf_counter()
cs, fixed, copies, adv, inst = tinyram_circuit.build(PL, be, seed=40, scale=scale)
torch.cuda.synchronize()
t_witness = time.perf_counter() - t0
assert cs.degree() == 6
inst_lists = None
if verify:
    L = int((inst[0].cpu().numpy().view(np.uint64).any(axis=1)).nonzero()[0].max()) + 1 if inst[0].any() else 1
    L = max(L, 1 << 12 if be.n > (1 << 13) else 1)
    inst_lists = [be._ints(c[:min(L, be.n - 6)].cpu().numpy().view(np.uint64)) for c in inst]
t0 = time.perf_counter()
pk = PL.keygen(be, cs, fixed, copies)
torch.cuda.synchronize()
t_keygen = time.perf_counter() - t0
rng = Rng(k)
launches0 = ctx.launches
torch.cuda.reset_peak_memory_stats()
t0 = time.perf_counter()
phases = {}
proof = PL.create_proof(be, pk, inst, adv, rng, PL.Blake2bWrite(C.base.p, p), debug="--debug" in sys.argv, timings=phases)
torch.cuda.synchronize()
t_prove = time.perf_counter() - t0
res = {"k": k, "scale": scale, "advice": cs.num_advice, "instance": cs.num_instance, "fixed": cs.num_fixed, "gates": len(cs.gates),
       "lookups": len(cs.lookups), "equality_columns": len(cs.permutation), "cs_degree": cs.degree(), "proof_bytes": len(proof),
       "params_new_s": round(t_params, 3), "witness_s": round(t_witness, 3), "keygen_s": round(t_keygen, 3),
       "create_proof_s": round(t_prove, 3), "phases_s": {k_: round(v, 3) for k_, v in phases.items()}, "kernel_launches_in_create_proof": ctx.launches - launches0,
       "torch_peak_gib": round(torch.cuda.max_memory_allocated() / 2**30, 1)}
if verify:
    bf, sf = O.BASE_FIELD[O.VESTA], O.SCALAR_FIELD[O.VESTA]
    g_l = np.concatenate([np.asarray(be.params.g_lagrange_points).reshape(-1, 8), np.asarray(be.params.w).reshape(1, 8)])
    g_c = np.asarray(be.params.g_points).reshape(-1, 8)
    as_pts = lambda arr: [None if not r.any() else tuple(be._ints(r.reshape(2, 4), be.q, be.Rqinv)) for r in np.asarray(arr).reshape(-1, 8)]

    class Marker(list):
        """a list of points that remembers the limb array it was made from (so the big MSMs skip the int -> limb conversion)"""
        limbs = None
        def __add__(self, other):
            out = Marker(list.__add__(self, other)); out.limbs = self.limbs; return out

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

    FC = FastCurve("vesta", pm.Fq, pm.Fp)
    gl = Marker(as_pts(g_l[:-1])); gl.limbs = g_l
    gc = Marker(as_pts(g_c)); gc.limbs = g_c
    params = {"k": k, "n": be.n, "g": gc, "g_lagrange": gl, "w": as_pts(be.params.w)[0], "u": as_pts(be.params.u)[0]}
    t0 = time.perf_counter()
    res["verified"] = VM.verify_proof(FC, params, pk.vk, inst_lists, proof)
    res["verify_s"] = round(time.perf_counter() - t0, 2)
    res["verify_error"] = None if res["verified"] else VM.verify_proof.last_error
    bad = bytearray(proof); bad[len(proof) // 3] ^= 2
    res["tampered_rejected"] = not VM.verify_proof(FC, params, pk.vk, inst_lists, bytes(bad))
print(json.dumps(res))
sys.exit(0 if (not verify or (res["verified"] and res["tampered_rejected"])) else 1)
