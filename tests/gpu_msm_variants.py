"""Ad-hoc (not a test): A/B of the MSM kernel variants on one B200 (results: profiles/msm_variants_r02.md).

Each variant runs in its own process (the switches are read once per process): B columns x (2^K + 1) Vesta points, device
resident (bench.py's step), 5 timed steps with CUDA events; the normalised results must be byte-identical across variants.

  python tests/gpu_msm_variants.py                  # K = 20, B = 8, uniform scalars
  K=22 B=4 python tests/gpu_msm_variants.py
  SHAPE=tinyram python tests/gpu_msm_variants.py    # 90 % {0,1}, 8 % < 2^32, 2 % uniform
  SHAPE=sparse16 python tests/gpu_msm_variants.py   # what a real advice column looks like: tinyram-shaped values on the first 2^16
                                                    # rows (the execution table), zero elsewhere, 6 uniform blinding rows at the end

Switches: TRP_MSM_C=c (window
width), TRP_MSM_REDUCE=1/2 (one doubling chain per chunk / the two-level weighted bucket sum of csrc/bucket_reduce.cuh)."""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = ([("default", {})]
            + [(f"c = {c}", {"TRP_MSM_C": str(c)}) for c in (13, 14, 15, 16, 17, 18, 19, 20, 21)]
            + [(f"c = {c}, chunk-chain reduction", {"TRP_MSM_C": str(c), "TRP_MSM_REDUCE": "1"}) for c in (18,)]
            + [(f"c = {c}, two-level reduction", {"TRP_MSM_C": str(c), "TRP_MSM_REDUCE": "2"}) for c in (16, 17)])


def make_scalars(n, m, shape, seed=20):
    """(m, n, 4) uint64 limbs + whether they are canonical small values that still need the Montgomery factor"""
    import numpy as np
    from tiny_ram_halo2_b200 import synthetic
    if shape == "uniform":
        return synthetic.random_scalars(n, seed, m), False
    rng = np.random.Generator(np.random.PCG64(seed))
    rows = n if shape == "tinyram" else (1 << 16)
    a = np.zeros((m, n, 4), dtype=np.uint64)
    kind = rng.random((m, rows))
    a[:, :rows, 0] = np.where(kind < 0.9, rng.integers(0, 2, size=(m, rows), dtype=np.uint64),
                              rng.integers(0, 1 << 32, size=(m, rows), dtype=np.uint64))
    wide = kind >= 0.98
    u = rng.integers(0, 1 << 64, size=(m, rows, 4), dtype=np.uint64)
    u[..., 3] &= np.uint64((1 << 62) - 1)
    a[:, :rows][wide] = u[wide]
    if shape == "sparse16":
        tail = rng.integers(0, 1 << 64, size=(m, 7, 4), dtype=np.uint64)       # blinding rows + the blind's slot
        tail[..., 3] &= np.uint64((1 << 62) - 1)
        a[:, n - 7:] = tail
    return a, True


def child():
    sys.path.insert(0, ROOT)
    import ctypes
    import numpy as np
    import torch
    import __graft_entry__ as ge
    pkg = ge.load_package()
    from tiny_ram_halo2_b200 import synthetic
    k = int(os.environ.get("K", "20"))
    n, m = (1 << k) + 1, int(os.environ.get("B", "8"))
    ctx = pkg.Context(0, pkg.VESTA)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    lib = ctx.lib
    d_pts = torch.empty((n, 8), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    synthetic.device_points(ctx, n, d_pts.data_ptr())
    hb = ctypes.c_void_p()
    ctx.check(lib.trp_dev_bases_load(ctx.handle, d_pts.data_ptr(), n, ctypes.byref(hb)))
    host, canonical = make_scalars(n, m, os.environ.get("SHAPE", "uniform"))
    d_scalars = torch.from_numpy(host.view(np.int64)).cuda()
    if canonical:                                   # canonical values -> Montgomery form: multiply by R^2 on the device
        R2 = np.array([(pow(2, 512, 0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001) >> (64 * i)) & (2**64 - 1)
                       for i in range(4)], dtype=np.uint64)
        d_r2 = torch.from_numpy(R2.view(np.int64)).cuda()
        torch.cuda.synchronize()
        ctx.check(lib.trp_dev_field_op(ctx.handle, 0, 2 | 16, d_scalars.data_ptr(), d_r2.data_ptr(), d_scalars.data_ptr(), m * n))
        ctx.sync()
    d_out = torch.zeros((m, 12), dtype=torch.int64, device="cuda")
    ctx.prof_reset(); ctx.prof_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(8):
        if i == 3:
            ctx.prof_reset()
            e0.record(stream)
        ctx.check(lib.trp_dev_msm_batch(ctx.handle, hb, d_scalars.data_ptr(), n, m, d_out.data_ptr()))
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    prof = {name: v[0] / max(v[1], 1) for name, v in ctx.prof_get().items() if v[1]}
    print(json.dumps({"ms_per_step": ms, "mpts": m * n / ms / 1e3, "phase_ms_per_launch": prof,
                      "sha256": hashlib.sha256(d_out.cpu().numpy().tobytes()).hexdigest()}))


def main():
    results = []
    for name, env in VARIANTS:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env={**os.environ, **env}, capture_output=True, text=True)
        if out.returncode:
            print(name, "FAILED", out.stderr[-2000:])
            continue
        r = json.loads(out.stdout.strip().splitlines()[-1])
        results.append((name, r))
        ph = r["phase_ms_per_launch"]
        print(f"{name:32s} {r['ms_per_step']:8.3f} ms/step {r['mpts']:7.1f} Mpts/s  sort {ph.get('msm_sort', float('nan')):.3f}  accumulate "
              f"{ph.get('msm_accum_l1', float('nan')):.3f}  levels {ph.get('msm_levels', float('nan')):.3f}  reduce {ph.get('msm_reduce', float('nan')):.3f} ms  {r['sha256'][:16]}")
    if len({r["sha256"] for _, r in results}) > 1:
        print("MISMATCH between variants")
        sys.exit(1)
    print("all variants byte-identical")


if __name__ == "__main__":
    child() if "--child" in sys.argv else main()
