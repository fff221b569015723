"""Ad-hoc (not a test): per-operation device times of one IPA round at half = 2^(K-1) (env K, default 20)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as ge
pkg = ge.load_package()
from tiny_ram_halo2_b200 import synthetic
from tiny_ram_halo2_b200._lib import ptr
K = int(os.environ.get("K", "20")); n = 1 << K; half = n // 2
ctx = pkg.Context(0, pkg.VESTA); lib = ctx.lib
stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
g = torch.empty((n, 8), dtype=torch.int64, device="cuda"); torch.cuda.synchronize()
synthetic.device_points(ctx, n, g.data_ptr()); ctx.sync()
a = torch.from_numpy(synthetic.random_scalars(n, 1).view(np.int64)).cuda()
b = torch.from_numpy(synthetic.random_scalars(n, 2).view(np.int64)).cuda()
out = torch.zeros((4, 12), dtype=torch.int64, device="cuda")
u = synthetic.random_scalars(1, 3)[0]
torch.cuda.synchronize()
def timed(name, fn, reps=3):
    fn(); ctx.sync()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); torch.cuda.synchronize()
    print(json.dumps({"op": name, "ms": round(e0.elapsed_time(e1) / reps, 4)}), flush=True)
for h in (half, half >> 4, half >> 8, half >> 12, 8):
    print("half", h)
    timed("msm_var", lambda: ctx.check(lib.trp_dev_msm_var(ctx.handle, g.data_ptr(), a.data_ptr(), h, 1, out.data_ptr())))
    timed("inner_product", lambda: ctx.check(lib.trp_dev_inner_products(ctx.handle, 0, a.data_ptr(), 0, b.data_ptr(), 0, h, 1, out.data_ptr())))
    timed("fold", lambda: ctx.check(lib.trp_dev_fold(ctx.handle, 0, a.data_ptr(), h, ptr(u))))
    gg = g.clone(); torch.cuda.synchronize()
    timed("collapse", lambda: ctx.check(lib.trp_dev_generator_collapse(ctx.handle, gg.data_ptr(), h, ptr(u))))
