"""Ad-hoc (not a test): a few batched NTTs of 8 x 2^K (env K, default 20) for profiling."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
pkg = ge.load_package()
from tiny_ram_halo2_b200._lib import ptr
K = int(os.environ.get("K", "20")); B = int(os.environ.get("B", "8"))
ctx = pkg.Context(0, pkg.VESTA)
stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
dom = pkg.EvaluationDomain(ctx, 6, K)
a = torch.randint(0, 1 << 62, (B, 1 << K, 4), dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
for i in range(4):
    if i == 1: e0.record(stream)
    ctx.check(ctx.lib.trp_dev_ntt(ctx.handle, a.data_ptr(), B, K, ptr(dom.omega)))
e1.record(stream); torch.cuda.synchronize()
print("ntt ms", e0.elapsed_time(e1) / 3)
