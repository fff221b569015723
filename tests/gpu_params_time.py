"""Times Params::new(k) on the device (row f3): hash-to-curve of the 2^k generators and the group iFFT to the Lagrange basis.
usage: python tests/gpu_params_time.py [k ...]"""
import ctypes
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import __graft_entry__ as ge

pkg = ge.load_package()
ctx = pkg.Context(0, pkg.VESTA)
zero = (ctypes.c_uint8 * 1)(0)
for k in [int(a) for a in sys.argv[1:]] or [12, 16, 18, 20]:
    n = 1 << k
    g = torch.zeros((n, 8), dtype=torch.int64, device="cuda")
    gl = torch.zeros((n, 8), dtype=torch.int64, device="cuda")
    wu = torch.zeros((2, 8), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    for rep in range(2):
        t0 = time.perf_counter()
        ctx.check(ctx.lib.trp_dev_hash_to_curve(ctx.handle, b"Halo2-Parameters", zero, 1, 1, 0, n, g.data_ptr()))
        ctx.sync()
        t1 = time.perf_counter()
        ctx.check(ctx.lib.trp_dev_params_new(ctx.handle, k, g.data_ptr(), gl.data_ptr(), wu.data_ptr()))
        ctx.sync()
        t2 = time.perf_counter()
    print(json.dumps({"k": k, "hash_to_curve_s": round(t1 - t0, 4), "params_new_s": round(t2 - t1, 4),
                      "group_fft_scalar_muls": (n // 2) * k + n}))
