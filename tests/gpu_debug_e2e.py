"""Ad-hoc GPU check (not a test): device-resident batch MSM vs host-pointer batch MSM vs the oracle at bench size."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as ge
pkg = ge.load_package()
from tiny_ram_halo2_b200 import synthetic
from util import O

K = int(os.environ.get("K", "20"))
n, m = (1 << K) + 1, int(os.environ.get("M", "8"))
ctx = pkg.Context(0, pkg.VESTA)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
lib = ctx.lib
d_pts = torch.empty((n, 8), dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
synthetic.device_points(ctx, n, d_pts.data_ptr())
ctx.sync()
hb = ctypes.c_void_p()
ctx.check(lib.trp_dev_bases_load(ctx.handle, d_pts.data_ptr(), n, ctypes.byref(hb)))
ctx.sync()
h_sc = torch.from_numpy(synthetic.random_scalars(n, 20, m).view(np.int64)).pin_memory()
d_sc = h_sc.cuda()
torch.cuda.synchronize()
outs = []
for rep in range(3):
    d_out = torch.zeros((m, 12), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctx.check(lib.trp_dev_msm_batch(ctx.handle, hb, d_sc.data_ptr(), n, m, d_out.data_ptr()))
    ctx.sync()
    outs.append(d_out.cpu().numpy().view(np.uint64).copy())
for rep in range(2):
    h_out = torch.zeros((m, 12), dtype=torch.int64).pin_memory()
    ctx.check(lib.trp_msm_batch(ctx.handle, hb, h_sc.data_ptr(), n, m, h_out.data_ptr()))
    outs.append(h_out.numpy().view(np.uint64).copy())
names = ["dev0", "dev1", "dev2", "host0", "host1"]
for i in range(1, len(outs)):
    eq = [bool(np.array_equal(outs[0][k], outs[i][k])) for k in range(m)]
    print(f"{names[0]} vs {names[i]}: {eq}", flush=True)
# one column at a time on the device path
single = []
for k in range(m):
    d_o = torch.zeros((1, 12), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctx.check(lib.trp_dev_msm_batch(ctx.handle, hb, d_sc[k].data_ptr(), n, 1, d_o.data_ptr()))
    ctx.sync()
    single.append(d_o.cpu().numpy().view(np.uint64)[0].copy())
print("dev0 vs dev-single:", [bool(np.array_equal(outs[0][k], single[k])) for k in range(m)], flush=True)
print("host0 vs dev-single:", [bool(np.array_equal(outs[3][k], single[k])) for k in range(m)], flush=True)
# oracle for column 0 and the last column
pts = d_pts.cpu().numpy().view(np.uint64)
sc = h_sc.numpy().view(np.uint64)
for k in (0, m - 1):
    t = time.time()
    want = O.msm(O.VESTA, sc[k], pts)
    dt = time.time() - t
    for nm, o in (("dev0", outs[0]), ("host0", outs[3])):
        got = O.jacobian_to_affine(O.VESTA, o[k].reshape(3, 4))
        print(f"col {k}: {nm} vs oracle: {bool(np.array_equal(got, want))}  (oracle {dt:.2f}s)", flush=True)
print("z limbs dev0:", outs[0][:, 8:], flush=True)
