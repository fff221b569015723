"""Ad-hoc (not a test): step-by-step wall times of one IPA opening (ipa.create_proof(trace=...)) at k = K (default 20), on one GPU
or under torchrun on several (the sharded rounds).  usage: [torchrun ...] python tests/gpu_ipa_trace.py"""
import json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
import numpy as np
import torch
import torch.distributed as dist
import __graft_entry__ as ge
pkg = ge.load_package()
from tiny_ram_halo2_b200 import ipa, synthetic, plonk as PL
from tiny_ram_halo2_b200.sharded_backend import ShardedRng
K = int(os.environ.get("K", "20")); n = 1 << K
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
d = None
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local)); d = dist
ctx = pkg.Context(local, pkg.VESTA)
stream = ctx.bind_torch_stream()
g = torch.empty((n + 2, 8), dtype=torch.int64, device="cuda"); torch.cuda.synchronize()
synthetic.device_points(ctx, n + 2, g.data_ptr()); ctx.sync()
gh = g.cpu().numpy().view(np.uint64)
params = ipa.IpaParams(ctx, K, gh[:n], gh[n], gh[n + 1])
p = 0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001
poly = torch.from_numpy(synthetic.random_scalars(n, 5).view(np.int64)).cuda()


class T:
    def __init__(self): self.h = __import__("hashlib").blake2b(b"x")
    def write_point(self, l): self.h.update(np.asarray(l).tobytes())
    def write_scalar(self, s): self.h.update(int(s).to_bytes(32, "little"))
    def squeeze_challenge_scalar(self):
        v = int.from_bytes(self.h.digest(), "little") % p; self.h.update(b"c"); return v or 1


digests, totals = [], []
for rep in range(3):
    rng = ShardedRng(p, seed=b"\x09" * 32)
    rng.prefetch(n, 1)
    tr, trace = T(), ([] if rep == 2 else None)
    torch.cuda.synchronize(); t0 = __import__("time").perf_counter()
    ipa.create_proof(params, rng, tr, poly, 12345, 0x1234567, rand_vector=rng.vector, dist=d, trace=trace)
    torch.cuda.synchronize(); total = __import__("time").perf_counter() - t0
    digests.append(tr.h.hexdigest()); totals.append(total)
if rank == 0:
    steps = [(b[0], round((b[1] - a[1]) * 1e3, 2)) for a, b in zip(trace, trace[1:])]
    print(json.dumps({"k": K, "n_gpus": world, "total_ms_untraced": round(totals[1] * 1e3, 1), "traced_total_ms": round(total * 1e3, 1), "transcript": digests[-1][:16],
                      "same_transcript_every_rep": len(set(digests)) == 1, "steps_ms": steps}))
if d is not None:
    dist.barrier(); dist.destroy_process_group()
