"""plonk.create_proof of the reference's REAL TinyRamCircuit on several GPUs of one box (sharded_backend.ShardedGpuBackend:
commitments sharded by column, the quotient by coset, replicated transcript), under torchrun:
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/gpu_multi_tinyram.py [W] [k] [--check] [--verify]
--check   rank 0 also proves on its GPU alone and compares the proof bytes (needs room for two backends: k <= 18)
--verify  rank 0 runs the oracle's independent verifier on the proof"""
import hashlib
import json
import os
import random
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE); sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import __graft_entry__ as ge

args = [a for a in sys.argv[1:] if not a.startswith("--")]
W = int(args[0]) if args else 16
k = int(args[1]) if len(args) > 1 else 2 + W // 2
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
d = None
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d = dist
pkg = ge.load_package()
PL = pkg.plonk
from tiny_ram_halo2_b200 import programs, tinyram as TR
from tiny_ram_halo2_b200.sharded_backend import ShardedGpuBackend
import pasta_model as pm
C = pm.Vesta
p = C.scalar.p
ctx = pkg.Context(local, pkg.VESTA)


from tiny_ram_halo2_b200.sharded_backend import ShardedRng


def Rng(seed):                 # one AES-CTR stream on every rank (a fixed seed here: the runs are compared byte for byte)
    return ShardedRng(p, d, "cuda", seed=bytes([seed]) * 32)


tr = programs.longest_loop(W)
circ, fixed, copies, adv, inst = TR.build(PL, tr, k, dense=False)
cs = circ.cs


def prove(be, reps):
    d_fixed, d_adv, d_inst = TR.device_columns(be, fixed), TR.device_columns(be, adv), TR.device_columns(be, inst)
    t0 = time.perf_counter()
    pk = PL.keygen(be, cs, d_fixed, copies)
    torch.cuda.synchronize()
    t_keygen = time.perf_counter() - t0
    runs, proof = [], None
    for rep in range(reps):
        cols = [c.clone() for c in d_adv]
        phases = {}
        sharded = d is not None and getattr(be, "dist", None) is not None
        if sharded:                               # the one-GPU check run on rank 0 must not enter a collective
            d.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        proof = PL.create_proof(be, pk, d_inst, cols, Rng(7), PL.Blake2bWrite(C.base.p, p), timings=phases)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if sharded:
            d.all_reduce(dt, op=dist.ReduceOp.MAX)
        runs.append({"create_proof_s_max_over_ranks": round(float(dt.item()), 3), "phases_s": {k_: round(v, 3) for k_, v in phases.items()}})
    return pk, proof, t_keygen, runs


be = ShardedGpuBackend(ctx, k, cs.degree(), d)
pk, proof, t_keygen, runs = prove(be, 4)
digest = hashlib.sha256(proof).digest()
same = True
if d is not None:
    t = torch.tensor(list(digest), dtype=torch.uint8, device="cuda")
    parts = [torch.empty_like(t) for _ in range(world)]
    d.all_gather(parts, t)
    same = all(bool(torch.equal(parts[0], q)) for q in parts)
res = {"circuit": "TinyRamCircuit (tinyram.py)", "word_bits": W, "k": k, "n_gpus": world, "trace_steps": len(tr.exe), "proof_bytes": len(proof),
       "keygen_s": round(t_keygen, 3), "proofs": runs, "best_create_proof_s": min(r["create_proof_s_max_over_ranks"] for r in runs),
       "proof_identical_on_all_ranks": same, "torch_peak_gib": round(torch.cuda.max_memory_allocated() / 2**30, 1)}
if "--verify" in sys.argv and rank == 0:
    import verify_util as VU
    t0 = time.perf_counter()
    res["verified"], res["verify_error"] = VU.verify(be, pk.vk, inst, proof)
    res["verify_s"] = round(time.perf_counter() - t0, 2)
if "--pverify" in sys.argv:                 # the package's own verifier on the device, on every rank (replicated)
    from tiny_ram_halo2_b200 import verifier as V
    t0 = time.perf_counter()
    try:
        V.verify_proof(be, pk.vk, V.SingleVerifier(be), TR.device_columns(be, inst), V.Blake2bRead(proof, C.base.p, p))
        res["product_verifier_accepts"] = True
    except V.VerifyError as e:
        res["product_verifier_accepts"] = False; res["product_verifier_error"] = str(e)
    res["product_verify_s"] = round(time.perf_counter() - t0, 2)
if "--check" in sys.argv:
    be.close(); del be, pk
    torch.cuda.empty_cache()
    if rank == 0:
        single = PL.GpuBackend(ctx, k, cs.degree())
        _, proof1, _, runs1 = prove(single, 2)
        res["single_gpu_create_proof_s"] = min(r["create_proof_s_max_over_ranks"] for r in runs1)
        res["bit_exact_vs_single_gpu"] = proof1 == proof
    if d is not None:
        d.barrier()
if rank == 0:
    print(json.dumps(res))
if d is not None:
    d.destroy_process_group()
ok = res["proof_identical_on_all_ranks"] and res.get("verified", True) and res.get("bit_exact_vs_single_gpu", True) and res.get("product_verifier_accepts", True)
sys.exit(0 if ok else 1)
