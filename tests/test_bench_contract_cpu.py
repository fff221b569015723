"""CPU checks around bench.py: (1) without a CUDA device the product arm refuses to run (there is no CPU fallback -- the CPU
restatement is only reachable as the explicitly labelled reference arm); (2) the bench lines committed under profiles/ carry every
key of the bench contract and are internally consistent (value = 1 / seconds, roofline.frac = achieved / peak, clean clocks, the
in-band verification of the headline proof)."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--no-cpu", "--no-extras"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stdout + r.stderr)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]          # no JSON line: nothing was measured


def _lines():
    out = []
    for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "bench_r02_n*.json"))):
        with open(p) as f:
            rows = [json.loads(l) for l in f if l.startswith("{")]
        assert len(rows) == 1, p
        out.append((os.path.basename(p), rows[0]))
    return out


def test_committed_bench_lines_follow_the_contract():
    lines = _lines()
    assert {d["n_gpus"] for _, d in lines} == {1, 2, 4, 8}
    for name, d in lines:
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                    "dtype", "data", "config", "roofline", "cpu_baseline", "clocks", "e2e", "gpu_launches"):
            assert key in d, (name, key)
        assert d["metric"] == "tinyram_create_proof_throughput" and d["unit"] == "proofs/s" and d["higher_is_better"] is True
        assert d["scaling"] == "strong" and d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
        assert d["warmup"] >= 3 and abs(d["value"] * d["ms_per_step"] / 1e3 - 1) < 1e-9
        rf = d["roofline"]
        assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and 0.5 < rf["frac"] < 1 and rf["traffic"] > 0
        assert 0.4 < sum(rf["phase_share_of_step"].values()) <= 1.0
        e = d["e2e"]
        assert e["unit"] == "proofs/s" and e["h2d_bytes_per_step"] > 1e8 and e["d2h_bytes_per_step"] == d["proof_bytes"] == 42560
        assert abs(e["value"] * e["seconds"] - 1) < 1e-9 and e["value"] != d["value"]
        assert d["clocks"]["reasons"] == [] and d["clocks"]["sm_mhz"] >= 0.95 * d["clocks"]["sm_max_mhz"]
        assert d["gpu_launches"] > 1000 * d["steps"]
        assert d["verified"] is True and d["tampered_proof_rejected"] is True and d["proof_identical_on_all_ranks"] is True
        if d["n_gpus"] == 1 and d["cpu_baseline"] is not None:
            c = d["cpu_baseline"]
            assert c["kind"] == "port" and c["cores"] >= 1 and c["sample"] and c["unit"] == "proofs/s" and c["value"] < d["value"] / 100
    by_n = {}
    for name, d in lines:
        by_n.setdefault(d["n_gpus"], []).append(d["ms_per_step"])
    t = {n: min(v) for n, v in by_n.items()}
    assert t[1] > t[2] > t[4] > t[8] and 4.0 < t[1] / t[8] < 8.0          # the strong-scaling claim of DESIGN.md section 5
