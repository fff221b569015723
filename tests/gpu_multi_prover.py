"""Sharded create_proof hot-path model under torchrun (BASELINE.json configs[4]).
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/gpu_multi_prover.py --k 20 [--scale S] [--check]
--check also runs the same workload on rank 0 alone and compares commitments and h(X) bit for bit."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--streamed", action="store_true", help="blockwise exchange: no rank holds all coefficient columns (default from k = 21)")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    d = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        d = dist
    pkg = ge.load_package()
    from tiny_ram_halo2_b200.sharded_model import ShardedProverModel
    ctx = pkg.Context(local, pkg.VESTA)
    st = torch.cuda.Stream()
    ctx.set_stream(st.cuda_stream)
    model = ShardedProverModel(ctx, a.k, st, d, scale=a.scale, streamed=True if a.streamed else None)
    model.prove_once()
    best = None
    for _ in range(a.reps):
        if d is not None:
            d.barrier()
        torch.cuda.synchronize()
        t, commitments, h, h_commit = model.prove_once()
        tt = torch.tensor([t["total_ms"]], dtype=torch.float64, device="cuda")
        if d is not None:
            d.all_reduce(tt, op=dist.ReduceOp.MAX)
        t["total_ms_max_over_ranks"] = float(tt.item())
        if best is None or t["total_ms_max_over_ranks"] < best["total_ms_max_over_ranks"]:
            best = t
    res = {"k": a.k, "scale": a.scale, "n_gpus": world, "streamed": model.streamed and world > 1, "torch_peak_gib": round(torch.cuda.max_memory_allocated() / 2**30, 1), "per_proof_columns": model.n_proof, "cosets": model.cosets,
           "phases_ms_rank0": {k: round(v, 2) for k, v in best.items()}}
    if a.check:
        ok = None
        commitments, h, h_commit = commitments.cpu(), h.cpu(), h_commit.cpu()
        model.close(); del model
        torch.cuda.empty_cache()
        if rank == 0:
            single = ShardedProverModel(ctx, a.k, st, None, scale=a.scale)
            t1, c1, h1, hc1 = single.prove_once()
            ok = bool(torch.equal(c1.cpu(), commitments) and torch.equal(h1.cpu(), h) and torch.equal(hc1.cpu(), h_commit))
            res["single_gpu_total_ms"] = round(t1["total_ms"], 2)
            res["bit_exact_vs_single_gpu"] = ok
        if d is not None:
            d.barrier()
    if rank == 0:
        print(json.dumps(res))
    if d is not None:
        d.destroy_process_group()
    if a.check and rank == 0 and not res["bit_exact_vs_single_gpu"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
