"""Ad-hoc timing script for the GPU box: microbenchmarks + per-phase timing breakdown (not a test, not the bench)."""
import sys, time, os, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as ge
pkg = ge.load_package()
from tiny_ram_halo2_b200 import synthetic
from tiny_ram_halo2_b200._lib import ptr
ctx = pkg.Context(0, pkg.VESTA)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
res = {}
for kind, name in enumerate(["imad_wide", "imad32", "iadd3x", "madc_pairs", "fe_mul", "fe_addsub"]):
    res[name] = round(ctx.microbench(kind, 512), 1)
print("MICROBENCH G-ops/s:", json.dumps(res), flush=True)

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

sizes = [int(x) for x in os.environ.get("LOGNS", "16,20,22").split(",")]
for logn in sizes:
    n = (1 << logn) + 1
    d_pts = torch.empty((n, 8), dtype=torch.int64, device="cuda")
    synthetic.device_points(ctx, n, d_pts.data_ptr()); ctx.sync()
    for flags, nm in ((2, "precomp"), (1, "windows")):
        hb = ctypes.c_void_p()
        t = time.time()
        ctx.check(ctx.lib.trp_bases_load_ex(ctx.handle, d_pts.cpu().numpy().view(np.uint64).ctypes.data_as(ctypes.c_void_p), n, flags, ctypes.byref(hb)))
        ctx.sync(); tl = time.time() - t
        for shape in ("uniform", "tinyram"):
            sc = synthetic.random_scalars(n, 20)
            if shape == "tinyram":
                rng = np.random.Generator(np.random.PCG64(5)); kind = rng.random(n)
                small = np.zeros((n, 4), dtype=np.uint64); small[:, 0] = np.where(kind < 0.9, rng.integers(0, 2, n), rng.integers(0, 1 << 32, n)).astype(np.uint64)
                # canonical small values must be converted to Montgomery form: multiply by R via the library
                r2 = np.tile(np.array([0x8c78ecb30000000f, 0xd7d30dbd8b0de0e7, 0x7797a99bc3c95d18, 0x096d41af7b9cb714], dtype=np.uint64), (n, 1))
                small = ctx.field_op("mul", small, r2)
                sc = np.where((kind < 0.98)[:, None], small, sc)
            d_sc = torch.from_numpy(sc.view(np.int64)).cuda(); d_out = torch.zeros(12, dtype=torch.int64, device="cuda")
            fn = lambda: ctx.check(ctx.lib.trp_dev_msm_batch(ctx.handle, hb, d_sc.data_ptr(), n, 1, d_out.data_ptr()))
            ms = timeit(fn)
            ctx.prof_reset(); ctx.prof_enable(True)
            for _ in range(3): fn()
            prof = {k: round(v[0] / max(v[1], 1), 4) for k, v in ctx.prof_get().items() if v[1]}
            ctx.prof_enable(False)
            print(f"msm logn={logn} {nm} {shape}: {ms:.3f} ms = {n/ms/1e3:.1f} Mpts/s  load={tl:.2f}s phases(ms)={prof}", flush=True)
        ctx.lib.trp_bases_free(hb)
    del d_pts
for logn in (16, 20, 23, 24):
    for batch in (1, 8):
        N = 1 << logn
        a = torch.randint(0, 2**62, (batch, N, 4), dtype=torch.int64, device="cuda")
        om = np.zeros(4, dtype=np.uint64)
        dom = pkg.EvaluationDomain(ctx, 2, logn)
        om = dom.omega
        fn = lambda: ctx.check(ctx.lib.trp_dev_ntt(ctx.handle, a.data_ptr(), batch, logn, ptr(om)))
        ms = timeit(fn)
        print(f"ntt logn={logn} batch={batch}: {ms:.3f} ms = {batch*N*64/ms/1e6:.1f} GB/s algorithmic, {batch*N/2*logn*128/ms/1e9:.2f} TMAC/s", flush=True)
        del a
