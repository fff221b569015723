"""Ad-hoc (not a test): run the create_proof workload model (env K, default 20)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
pkg = ge.load_package()
from tiny_ram_halo2_b200.prover_model import CreateProofModel
K = int(os.environ.get("K", "20"))
ctx = pkg.Context(0, pkg.VESTA)
stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
t = time.time()
m = CreateProofModel(ctx, K, stream)
print("setup_s", round(time.time() - t, 2), json.dumps(m.describe()), flush=True)
for i in range(3):
    r = m.prove_once()
    print(json.dumps({k: round(v, 2) for k, v in r.items()}), flush=True)
print("max_mem_GB", round(torch.cuda.max_memory_allocated() / 2**30, 1))
