"""Ad-hoc first-contact script for the GPU box: microbenchmarks + timing breakdown (not a test, not the bench)."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from util import O, make_points, scalars_uniform, scalars_tinyram, affine_of
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package()
ctx = pkg.Context(0, pkg.VESTA)
res = {}
for kind, name in enumerate(["imad_wide", "imad32", "iadd3x", "madc_pairs", "fe_mul", "fe_addsub"]):
    res[name] = ctx.microbench(kind, 512)
print("MICROBENCH G-ops/s:", json.dumps(res))
import torch
for logn in (16, 20):
    n = (1 << logn) + 1
    t = time.time(); pts = make_points(O.VESTA, n); print("gen pts", time.time() - t)
    for flags in (2, 1):
        t = time.time(); bases = pkg.Bases(ctx, pts, flags); ctx.sync(); print(f"bases load flags={flags} n={n}", time.time() - t, bases.describe())
        for shape in ("uniform", "tinyram"):
            sc = scalars_uniform(O.VESTA, n) if shape == "uniform" else scalars_tinyram(O.VESTA, n)
            out = pkg.best_multiexp(ctx, sc, bases)
            times = []
            for _ in range(3):
                t = time.time(); out = pkg.best_multiexp(ctx, sc, bases); times.append(time.time() - t)
            print(f"msm e2e(host ptr) logn={logn} flags={flags} {shape}: {min(times)*1e3:.3f} ms")
            # device-resident timing
            d_sc = torch.from_numpy(sc.view(np.int64)).cuda(); d_out = torch.zeros(12, dtype=torch.int64, device="cuda")
            torch.cuda.synchronize()
            ctx.set_stream(torch.cuda.current_stream().cuda_stream)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            for _ in range(2):
                ctx.check(ctx.lib.trp_dev_msm_batch(ctx.handle, bases.handle, d_sc.data_ptr(), n, 1, d_out.data_ptr()))
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                ctx.check(ctx.lib.trp_dev_msm_batch(ctx.handle, bases.handle, d_sc.data_ptr(), n, 1, d_out.data_ptr()))
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"msm device logn={logn} flags={flags} {shape}: {ms:.3f} ms = {n/ms/1e3:.1f} Mpts/s")
            ctx.set_stream(0)
        bases.free()
# NTT timing
for logn in (16, 20, 23):
    for batch in (1, 8):
        N = 1 << logn
        a = torch.randint(0, 2**62, (batch, N, 4), dtype=torch.int64, device="cuda")
        a[..., 3] &= (1 << 61) - 1
        om = O.to_mont(O.FP, O.ints_to_limbs([pow(0x2bce74deac30ebda362120830561f81aea322bf2b7bb7584bdad6fabd87ea32f, 1 << (32 - logn), O.MODULUS[O.FP])]))[0]
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        from tiny_ram_halo2_b200._lib import ptr
        for _ in range(2):
            ctx.check(ctx.lib.trp_dev_ntt(ctx.handle, a.data_ptr(), batch, logn, ptr(om)))
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ctx.check(ctx.lib.trp_dev_ntt(ctx.handle, a.data_ptr(), batch, logn, ptr(om)))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"ntt logn={logn} batch={batch}: {ms:.3f} ms = {batch*N*64/ms/1e6:.1f} GB/s algorithmic")
        ctx.set_stream(0)
