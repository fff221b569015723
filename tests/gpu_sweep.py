"""Ad-hoc (not a test): BASELINE.json configs[1] and configs[2] -- the MSM sweep 2^16..2^24 (+1) on Vesta and Pallas, uniform
AND TinyRAM-shaped scalars, and the batched NTT / iNTT / coset sweep 2^16..2^24 x 8..64 columns over Fp.  Every case is first
COMPARED with the CPU oracle (column 0, byte for byte: BASELINE.md section 3 accepts a timing only then), then timed device
resident with CUDA events.  One JSON line per case.
  WHICH=msm|ntt  CHECK=0 (skip the oracle compare)  python tests/gpu_sweep.py"""
import ctypes, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
import numpy as np
import torch
import __graft_entry__ as ge
pkg = ge.load_package()
from tiny_ram_halo2_b200 import synthetic
from tiny_ram_halo2_b200._lib import ptr
from gpu_msm_variants import make_scalars
from util import O

CHECK = os.environ.get("CHECK", "1") != "0"
MODULUS = {pkg.VESTA: O.MODULUS[O.FP], pkg.PALLAS: O.MODULUS[O.FQ]}          # scalar fields


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
    for curve, cname, ocurve in ((pkg.VESTA, "vesta", O.VESTA), (pkg.PALLAS, "pallas", O.PALLAS)):
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
            for shape in (("uniform", "tinyram") if cname == "vesta" else ("uniform",)):
                host, canonical = make_scalars(n, m, shape, seed=20 + logn)
                sc = torch.from_numpy(host.view(np.int64)).cuda()
                if canonical:
                    p = MODULUS[curve]
                    r2 = np.array([(pow(2, 512, p) >> (64 * i)) & (2**64 - 1) for i in range(4)], dtype=np.uint64)
                    d_r2 = torch.from_numpy(r2.view(np.int64)).cuda(); torch.cuda.synchronize()
                    ctx.check(ctx.lib.trp_dev_field_op(ctx.handle, 0, 2 | 16, sc.data_ptr(), d_r2.data_ptr(), sc.data_ptr(), m * n)); ctx.sync()
                out = torch.zeros((m, 12), dtype=torch.int64, device="cuda"); torch.cuda.synchronize()
                run = lambda: ctx.check(ctx.lib.trp_dev_msm_batch(ctx.handle, hb, sc.data_ptr(), n, m, out.data_ptr()))
                checked = None
                if CHECK:
                    run(); ctx.sync()
                    got = O.jacobian_to_affine(ocurve, out[0].cpu().numpy().view(np.uint64).reshape(3, 4))
                    want = O.msm(ocurve, sc[0].cpu().numpy().view(np.uint64), d_pts.cpu().numpy().view(np.uint64))
                    checked = bool(np.array_equal(got, want))
                    if not checked:
                        print(json.dumps({"case": "msm", "curve": cname, "log_n": logn, "scalars": shape, "MISMATCH": True}), flush=True)
                        sys.exit(1)
                ms = timeit(stream, run, 3)
                print(json.dumps({"case": "msm", "curve": cname, "log_n": logn, "n": n, "columns": m, "scalars": shape, "c": desc[0], "windows": desc[1],
                                  "precomp": desc[2], "equals_oracle": checked, "ms_per_msm": round(ms / m, 4), "Mpts_per_s": round(m * n / ms / 1e3, 1)}), flush=True)
                del sc
            ctx.lib.trp_bases_free(hb); del d_pts
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
            if CHECK and batch == 8:
                host = a[0].cpu().numpy().view(np.uint64).copy()
                ctx.check(ctx.lib.trp_dev_ntt(ctx.handle, a.data_ptr(), batch, logn, ptr(dom.omega))); ctx.sync()
                res["equals_oracle"] = bool(np.array_equal(a[0].cpu().numpy().view(np.uint64), O.fft(O.FP, host, logn, np.ascontiguousarray(dom.omega))))
                if not res["equals_oracle"]:
                    print(json.dumps(res), flush=True); sys.exit(1)
            for name, fn in (("ntt", lambda: ctx.check(ctx.lib.trp_dev_ntt(ctx.handle, a.data_ptr(), batch, logn, ptr(dom.omega)))),
                             ("lagrange_to_coeff", lambda: ctx.check(ctx.lib.trp_dev_lagrange_to_coeff(dom.handle, a.data_ptr(), batch))),
                             ("coeff_to_coset", lambda: ctx.check(ctx.lib.trp_dev_coeff_to_coset(dom.handle, a.data_ptr(), a.data_ptr(), batch, 3)))):
                ms = timeit(stream, fn, 3)
                res[name] = {"ms_per_column": round(ms / batch, 4), "algorithmic_GBps": round(batch * N * 64 / ms / 1e6, 1),
                             "TMACps_model": round(batch * (N / 2) * logn * 128 / ms / 1e9, 2)}
            print(json.dumps(res), flush=True)
            del a
        dom.free()
