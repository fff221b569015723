"""A REAL plonk.create_proof of the reference's TinyRamCircuit (tiny-ram-halo2_b200/tinyram.py: the actual gates, lookups and
witness, not the stand-in of tinyram_circuit.py) for a long trace, on one B200, checked by the oracle's independent verifier.
BASELINE.json configs[3]: word size 32 (program / execution tables of 2^16 rows, up to 65 535 steps), k = 20.
usage: python tests/gpu_tinyram_real.py [W] [k] [steps] [--no-verify] [--empty-keygen] [--lists]"""
import json
import os
import random
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE); sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import __graft_entry__ as ge
import pasta_model as pm
import tinyram_programs as TP
import verify_util as VU

args = [a for a in sys.argv[1:] if not a.startswith("--")]
W = int(args[0]) if args else 8
k = int(args[1]) if len(args) > 1 else 2 + W // 2
steps = int(args[2]) if len(args) > 2 else (1 << (W // 2)) - 1
verify = "--no-verify" not in sys.argv
import torch
pkg = ge.load_package()
PL = pkg.plonk
from tiny_ram_halo2_b200 import tinyram as TR, trace as T
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
        a[:, 3] &= np.uint64((1 << 62) - 1)
        return a


from tiny_ram_halo2_b200 import programs
t0 = time.perf_counter()
tr = programs.longest_loop(W, steps)                       # 2 set-up steps, (len(body) + 3) per pass, 1 Answer
t_trace = time.perf_counter() - t0
t0 = time.perf_counter()
circ, fixed, copies, adv, inst = TR.build(PL, tr, k, dense=False, keygen_from_empty_circuit="--empty-keygen" in sys.argv,
                                          arrays="--lists" not in sys.argv)       # uint64 array columns: the upload path of bench.py
t_synth = time.perf_counter() - t0
cs = circ.cs
t0 = time.perf_counter()
be = PL.GpuBackend(ctx, k, cs.degree())
torch.cuda.synchronize()
t_params = time.perf_counter() - t0
t0 = time.perf_counter()
d_fixed, d_adv, d_inst = TR.device_columns(be, fixed), TR.device_columns(be, adv), TR.device_columns(be, inst)
torch.cuda.synchronize()
t_upload = time.perf_counter() - t0
t0 = time.perf_counter()
pk = PL.keygen(be, cs, d_fixed, copies)
torch.cuda.synchronize()
t_keygen = time.perf_counter() - t0
runs = []
for rep in range(4):
    rng = Rng(k + rep)
    launches0 = ctx.launches
    torch.cuda.reset_peak_memory_stats()
    phases = {}
    cols = [c.clone() for c in d_adv]                      # create_proof writes the blinding rows in place
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    proof = PL.create_proof(be, pk, d_inst, cols, rng, PL.Blake2bWrite(C.base.p, p), debug=(rep == 0 and "--debug" in sys.argv), timings=phases)
    torch.cuda.synchronize()
    runs.append({"create_proof_s": round(time.perf_counter() - t0, 3), "phases_s": {k_: round(v, 3) for k_, v in phases.items()},
                 "kernel_launches": ctx.launches - launches0, "torch_peak_gib": round(torch.cuda.max_memory_allocated() / 2**30, 1)})
res = {"circuit": "TinyRamCircuit (tinyram.py)", "word_bits": W, "k": k, "trace_steps": len(tr.exe), "program_lines": len(tr.prog),
       "table_len": circ.table_len, "advice": cs.num_advice, "instance": cs.num_instance, "fixed": cs.num_fixed, "gates": len(cs.gates),
       "lookups": len(cs.lookups), "equality_columns": len(cs.permutation), "copies": sum(getattr(c, "rows", 1) for c in copies), "cs_degree": cs.degree(),
       "proof_bytes": len(proof), "interpreter_s": round(t_trace, 3), "synthesize_s": round(t_synth, 3), "params_new_s": round(t_params, 3),
       "upload_s": round(t_upload, 3), "keygen_s": round(t_keygen, 3), "first_proof": runs[0], "second_proof": runs[1], "later_proofs_s": [r["create_proof_s"] for r in runs[2:]],
       "best_proof": min(runs, key=lambda r: r["create_proof_s"])}
if verify:
    inst = [[int(v) for v in col] for col in inst]         # the oracle's verifier wants Python ints
    t0 = time.perf_counter()
    res["verified"], res["verify_error"] = VU.verify(be, pk.vk, inst, proof)
    res["verify_s"] = round(time.perf_counter() - t0, 2)
    bad = bytearray(proof); bad[len(proof) // 3] ^= 2
    res["tampered_rejected"] = not VU.verify(be, pk.vk, inst, bytes(bad))[0]
try:        # the package's own verifier on the device (verifier.py), instance columns as device vectors: reported, not gating
    from tiny_ram_halo2_b200 import verifier as V
    torch.cuda.synchronize(); t0 = time.perf_counter()
    V.verify_proof(be, pk.vk, V.SingleVerifier(be), d_inst, V.Blake2bRead(proof, be.q, be.p))
    torch.cuda.synchronize()
    res["product_verifier"] = {"accepted": True, "seconds": round(time.perf_counter() - t0, 3)}
    bad = bytearray(proof); bad[len(proof) // 3] ^= 2
    try:
        V.verify_proof(be, pk.vk, V.SingleVerifier(be), d_inst, V.Blake2bRead(bytes(bad), be.q, be.p))
        res["product_verifier"]["tampered_rejected"] = False
    except V.VerifyError:
        res["product_verifier"]["tampered_rejected"] = True
except Exception as e:
    res["product_verifier"] = {"error": repr(e)}
print(json.dumps(res))
sys.exit(0 if (not verify or (res["verified"] and res["tampered_rejected"])) else 1)
