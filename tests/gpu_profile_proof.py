"""Ad-hoc (not a test): where the HOST time of one create_proof goes (cProfile; the device work is asynchronous, so waits show up
in the calls that read results back), and the library's per-phase device time (trp_prof_*).
usage: python tests/gpu_profile_proof.py [W] [k]"""
import cProfile
import io
import json
import os
import pstats
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE); sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import __graft_entry__ as ge
import torch

W = int(sys.argv[1]) if len(sys.argv) > 1 else 32
k = int(sys.argv[2]) if len(sys.argv) > 2 else 20
pkg = ge.load_package()
from tiny_ram_halo2_b200 import plonk as PL, programs, tinyram as TR
from tiny_ram_halo2_b200.sharded_backend import ShardedGpuBackend, ShardedRng
ctx = pkg.Context(0, pkg.VESTA)
tr = programs.longest_loop(W)
circ, fixed, copies, adv, inst = TR.build(PL, tr, k, dense=False, arrays=True)
cs = circ.cs
be = ShardedGpuBackend(ctx, k, cs.degree(), None)
d_fixed, d_adv, d_inst = TR.device_columns(be, fixed), TR.device_columns(be, adv), TR.device_columns(be, inst)
pk = PL.keygen(be, cs, d_fixed, copies)
rng = ShardedRng(be.p, seed=b"\x05" * 32)


def prove(timings=None):
    cols = [c.clone() for c in d_adv]
    return PL.create_proof(be, pk, d_inst, cols, rng, PL.Blake2bWrite(be.q, be.p), timings=timings)


import hashlib
first_sha = hashlib.sha256(prove()).hexdigest()          # fixed seed: comparable between runs (e.g. TRP_COMMIT_BY_PARTS=0 / 1)
prove()
be._wait()
t0 = time.perf_counter(); prove(); be._wait(); wall = time.perf_counter() - t0
ph = {}
t0 = time.perf_counter(); prove(ph); be._wait(); wall_sync = time.perf_counter() - t0
ctx.prof_reset(); ctx.prof_enable(True)
prove(); be._wait()
prof, work = ctx.prof_get(), ctx.prof_work()
ctx.prof_enable(False)
pr = cProfile.Profile()
pr.enable(); prove(); be._wait(); pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(json.dumps({"k": k, "first_proof_sha256": first_sha, "create_proof_s": round(wall, 3), "with_phase_waits_s": round(wall_sync, 3), "phases_s": {a: round(b, 3) for a, b in ph.items()},
                  "device_ms_by_library_phase": {a: round(b[0], 1) for a, b in prof.items() if b[1]}, "spans": {a: b[1] for a, b in prof.items() if b[1]},
                  "ntt_butterflies": work["ntt_pass"]}))
print(s.getvalue()[:9000])
