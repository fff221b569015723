"""A real create_proof through plonk.py's GpuBackend at a larger k, verified by the oracle's independent verify_proof
(its MSMs run on the C++ oracle for speed).  usage: python tests/gpu_plonk_time.py [k]"""
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
import plonk_circuits

k = int(sys.argv[1]) if len(sys.argv) > 1 else 14
pkg = ge.load_package()
PL = pkg.plonk
C = pm.Vesta
p = C.scalar.p
ctx = pkg.Context(0, pkg.VESTA)
cs, fixed, copies, adv, inst = plonk_circuits.standard(PL)
t0 = time.perf_counter()
be = PL.GpuBackend(ctx, k, cs.degree())
t_params = time.perf_counter() - t0
usable = be.n - (cs.blinding_factors() + 1)
rnd = random.Random(k)
for r in range(4, usable):
    a, b = rnd.randrange(8), rnd.randrange(p)
    add = rnd.random() < 0.5
    adv[0].append(a); adv[1].append(b); adv[2].append((a + b) % p if add else a * b % p)
    fixed[0].append(int(add)); fixed[1].append(int(not add)); fixed[2].append(0); fixed[4].append(1)
t0 = time.perf_counter()
pk = PL.keygen(be, cs, fixed, copies)
t_keygen = time.perf_counter() - t0
launches0 = ctx.launches
t0 = time.perf_counter()
proof = PL.create_proof(be, pk, inst, adv, lambda: rnd.randrange(p), PL.Blake2bWrite(C.base.p, p))
t_prove = time.perf_counter() - t0
launches = ctx.launches - launches0


# the verifier's large MSMs on the C++ oracle (multithreaded); everything else stays the Python restatement
class FastCurve(pm.Curve):
    def best_multiexp(self, scalars, bases):
        if len(bases) < 256:
            return super().best_multiexp(scalars, bases)
        bf, sf = O.BASE_FIELD[O.VESTA], O.SCALAR_FIELD[O.VESTA]
        pts = np.zeros((len(bases), 8), dtype=np.uint64)
        flat = [c for P_ in bases for c in (P_ if P_ is not None else (0, 0))]
        pts[:] = O.to_mont(bf, O.ints_to_limbs(flat)).reshape(-1, 8)
        sc = O.to_mont(sf, O.ints_to_limbs([s % p for s in scalars]))
        out = O.msm(O.VESTA, sc, pts)
        if not out.any():
            return None
        x, y = O.limbs_to_ints(O.from_mont(bf, out.reshape(2, 4)))
        return (x, y)


FC = FastCurve("vesta", pm.Fq, pm.Fp)
prm = be.params
as_pts = lambda arr: [None if not r.any() else tuple(be._ints(r.reshape(2, 4), be.q, be.Rqinv)) for r in np.asarray(arr).reshape(-1, 8)]
params = {"k": k, "n": be.n, "g": as_pts(prm.g_points), "g_lagrange": as_pts(prm.g_lagrange_points), "w": as_pts(prm.w)[0], "u": as_pts(prm.u)[0]}
t0 = time.perf_counter()
ok = VM.verify_proof(FC, params, pk.vk, inst, proof)
bad = bytearray(proof); bad[200] ^= 1
rejected = not VM.verify_proof(FC, params, pk.vk, inst, bytes(bad))
t_verify = time.perf_counter() - t0
print(json.dumps({"k": k, "rows_used": usable, "proof_bytes": len(proof), "verified": ok, "tampered_rejected": rejected,
                  "params_new_s": round(t_params, 3), "keygen_s": round(t_keygen, 3), "create_proof_s": round(t_prove, 3),
                  "verify_two_proofs_s": round(t_verify, 3), "kernel_launches_in_create_proof": launches}))
sys.exit(0 if ok and rejected else 1)
