"""Ad-hoc (not a test): BASELINE.json configs 2 and 3 -- MSM sweep 2^16..2^24 (+1) on Vesta and Pallas, batched NTT / iNTT /
coset sweep 2^16..2^24 x 8..64 columns over Fp.  Prints one JSON line per case (device-resident, CUDA-event timed)."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as ge
pkg = ge.load_package()
from tiny_ram_halo2_b200 import synthetic
from tiny_ram_halo2_b200._lib import ptr

def timeit(stream, fn, reps):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

which = os.environ.get("WHICH", "msm,ntt").split(",")
if "msm" in which:
    for curve, cname in ((pkg.VESTA, "vesta"), (pkg.PALLAS, "pallas")):
        ctx = pkg.Context(0, curve); stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
        for logn in (16, 18, 20, 22, 24):
            if cname == "pallas" and logn not in (16, 20, 24): continue
            n = (1 << logn) + 1
            m = 8 if logn <= 20 else (4 if logn == 22 else 1)
            d_pts = torch.empty((n, 8), dtype=torch.int64, device="cuda"); torch.cuda.synchronize()
            synthetic.device_points(ctx, n, d_pts.data_ptr()); ctx.sync()
            hb = ctypes.c_void_p()
            ctx.check(ctx.lib.trp_dev_bases_load(ctx.handle, d_pts.data_ptr(), n, ctypes.byref(hb))); ctx.sync()
            desc = (ctypes.c_uint * 3)(); ctx.lib.trp_bases_describe(hb, desc)
            sc = torch.randint(0, 1 << 62, (m, n, 4), dtype=torch.int64, device="cuda")
            out = torch.zeros((m, 12), dtype=torch.int64, device="cuda"); torch.cuda.synchronize()
            ms = timeit(stream, lambda: ctx.check(ctx.lib.trp_dev_msm_batch(ctx.handle, hb, sc.data_ptr(), n, m, out.data_ptr())), 3)
            print(json.dumps({"case": "msm", "curve": cname, "log_n": logn, "n": n, "columns": m, "c": desc[0], "windows": desc[1], "precomp": desc[2],
                              "ms_per_msm": round(ms / m, 4), "Mpts_per_s": round(m * n / ms / 1e3, 1)}), flush=True)
            ctx.lib.trp_bases_free(hb); del d_pts, sc
        ctx.close()
if "ntt" in which:
    ctx = pkg.Context(0, pkg.VESTA); stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
    for logn in (16, 18, 20, 22, 24):
        dom = pkg.EvaluationDomain(ctx, 6, logn)
        N = 1 << logn
        for batch in (8, 64):
            if logn == 24 and batch == 64 and os.environ.get("SKIP_BIG"): continue
            a = torch.randint(0, 1 << 62, (batch, N, 4), dtype=torch.int64, device="cuda"); torch.cuda.synchronize()
            res = {"case": "ntt", "log_n": logn, "columns": batch}
            for name, fn in (("ntt", lambda: ctx.check(ctx.lib.trp_dev_ntt(ctx.handle, a.data_ptr(), batch, logn, ptr(dom.omega)))),
                             ("lagrange_to_coeff", lambda: ctx.check(ctx.lib.trp_dev_lagrange_to_coeff(dom.handle, a.data_ptr(), batch))),
                             ("coeff_to_coset", lambda: ctx.check(ctx.lib.trp_dev_coeff_to_coset(dom.handle, a.data_ptr(), a.data_ptr(), batch, 3)))):
                ms = timeit(stream, fn, 3)
                res[name] = {"ms_per_column": round(ms / batch, 4), "algorithmic_GBps": round(batch * N * 64 / ms / 1e6, 1),
                             "TMACps_model": round(batch * (N / 2) * logn * 128 / ms / 1e9, 2)}
            print(json.dumps(res), flush=True)
            del a
        dom.free()
