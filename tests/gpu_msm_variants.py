"""Ad-hoc (not a test): A/B of the opt-in MSM kernel variants against the default path on one B200.

Each variant runs in its own process (the switches are read once per process): 8 columns x (2^K + 1) Vesta points, uniform
scalars, device resident (bench.py's step), 5 timed steps with CUDA events; the normalised results must be byte-identical.

  python tests/gpu_msm_variants.py            # K = 20
  K=16 python tests/gpu_msm_variants.py

Variants: TRP_MSM_CALL=1 (field multiplication out of line in the level-1 accumulation), TRP_MSM_C=17..20 (wider windows: fewer
bucket additions per scalar, more buckets to reduce) and TRP_MSM_REDUCE=2 (the two-level weighted bucket sum of
csrc/bucket_reduce.cuh, which is what makes the wider windows affordable) and TRP_MSM_SEG=1 (level-1 tasks as aligned windows
of the sorted entry list: no partial tasks, whatever the bucket sizes) -- DESIGN.md section 9.  K=22 / K=24 are where wider windows
should pay most (set B=4 / B=1 columns to stay inside the scratch budget)."""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = ([("default", {}), ("fe_mul out of line", {"TRP_MSM_CALL": "1"}), ("two-level reduction", {"TRP_MSM_REDUCE": "2"})]
            + [(f"c = {c}", {"TRP_MSM_C": str(c)}) for c in (17, 18, 19, 20)]
            + [(f"c = {c}, two-level reduction", {"TRP_MSM_C": str(c), "TRP_MSM_REDUCE": "2"}) for c in (18, 19, 20)]
            + [("segmented level 1", {"TRP_MSM_SEG": "1"})]
            + [(f"c = {c}, segmented, two-level", {"TRP_MSM_C": str(c), "TRP_MSM_SEG": "1", "TRP_MSM_REDUCE": "2"}) for c in (18, 19, 20)])


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
    d_scalars = torch.from_numpy(synthetic.random_scalars(n, 20, m).view(np.int64)).cuda()
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
