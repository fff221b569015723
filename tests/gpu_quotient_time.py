"""Ad-hoc timing (not a test): quotient VM on one size-n coset of the TinyRAM-shaped program (env K, default 20)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as ge
pkg = ge.load_package()
from tiny_ram_halo2_b200 import tinyram_shape as T, poly as P
K = int(os.environ.get("K", "20"))
ctx = pkg.Context(0, pkg.VESTA)
stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
dom = pkg.EvaluationDomain(ctx, 6, K)
n = 1 << K
sh = T.build()
ev = P.new_evaluator(ctx)
prog = P.compile_ast(sh.ast, ev.modulus)
cols = torch.randint(0, 1 << 62, (sh.n_columns, n, 4), dtype=torch.int64, device="cuda")
out = torch.zeros((n << 3, 4), dtype=torch.int64, device="cuda")
ptrs = [cols[c].data_ptr() for c in range(sh.n_columns)]
torch.cuda.synchronize()
def run(cs): ev.evaluate_device(prog, dom, ptrs, out.data_ptr(), coset=cs)
run(0); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(stream)
for cs in range(1, 4): run(cs)
e1.record(stream); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
c = prog.counts()
print(json.dumps({"K": K, "ms_per_coset": round(ms, 3), "ms_all_8_cosets": round(8 * ms, 2), "counts": c, "n_regs": prog.n_regs,
                  "G_fmul_per_s": round(c["mul"] * n / ms / 1e6, 1), "load_GBps": round(c["load"] * n * 32 / ms / 1e6, 1)}))
