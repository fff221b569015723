"""Ad-hoc timing (not a test): one MSM of 2^K+1 points, per-phase device times.  env: K, M, REPS."""
import ctypes, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as ge
pkg = ge.load_package()
from tiny_ram_halo2_b200 import synthetic
K = int(os.environ.get("K", "20")); m = int(os.environ.get("M", "1")); reps = int(os.environ.get("REPS", "5"))
n = (1 << K) + 1
ctx = pkg.Context(0, pkg.VESTA)
stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
lib = ctx.lib
d_pts = torch.empty((n, 8), dtype=torch.int64, device="cuda"); torch.cuda.synchronize()
synthetic.device_points(ctx, n, d_pts.data_ptr()); ctx.sync()
hb = ctypes.c_void_p()
ctx.check(lib.trp_dev_bases_load(ctx.handle, d_pts.data_ptr(), n, ctypes.byref(hb))); ctx.sync()
d_sc = torch.from_numpy(synthetic.random_scalars(n, 20, m).view(np.int64)).cuda()
if os.environ.get("DIST", "") == "tinyram":      # 90 % {0,1}, 8 % < 2^32, 2 % uniform, Montgomery form
    from util import O
    kind = torch.rand((m, n), device="cuda")
    small = torch.zeros((m, n, 4), dtype=torch.int64, device="cuda")
    small[..., 0] = torch.where(kind < 0.9, torch.randint(0, 2, (m, n), device="cuda"), torch.randint(0, 1 << 32, (m, n), device="cuda"))
    r2 = torch.from_numpy(O.to_mont(O.FP, O.ints_to_limbs([(1 << 256) % O.MODULUS[O.FP]])).view(np.int64)).cuda()
    torch.cuda.synchronize()
    ctx.check(lib.trp_dev_field_op(ctx.handle, 0, 2 | 16, small.data_ptr(), r2.data_ptr(), small.data_ptr(), m * n)); ctx.sync()
    d_sc = torch.where((kind < 0.98)[..., None], small, d_sc).contiguous()
d_out = torch.zeros((m, 12), dtype=torch.int64, device="cuda"); torch.cuda.synchronize()
fn = lambda: ctx.check(lib.trp_dev_msm_batch(ctx.handle, hb, d_sc.data_ptr(), n, m, d_out.data_ptr()))
for _ in range(3): fn()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(reps): fn()
e1.record(stream); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
ctx.prof_reset(); ctx.prof_enable(True)
for _ in range(reps): fn()
prof = {k: round(v[0] / reps, 4) for k, v in ctx.prof_get().items() if v[1]}
ctx.prof_enable(False)
chk = int(d_out.cpu().numpy().view(np.uint64).sum() & 0xffffffff)
print(json.dumps({"tag": os.environ.get("TAG", ""), "K": K, "m": m, "ms": round(ms, 4), "Mpts/s": round(m * n / ms / 1e3, 1), "phases_ms": prof, "chk": chk}), flush=True)
