"""plonk.create_proof on several GPUs of one box (BASELINE.json configs[4]): one process per GPU, every process
runs the SAME host logic on the same witness with the same RNG, so the transcript is replicated and no rank waits for
another's challenges; the two heavy, naturally sharded parts of the path are divided:

  * commitments (best_multiexp): the columns of every commit_lagrange_many / commit_many batch go round-robin over the ranks
    (parallel.shard_columns), the 64-byte affine results are all_gathered -- advice / lookup / grand-product / h-piece / key
    commitments;
  * the quotient: rank r evaluates cosets r, r + G, ... of the j - 1 that determine h(X) (coset NTT of every per-proof
    polynomial + the quotient program), the n-value results are all_gathered (parallel.all_gather_columns) and every rank
    recovers h(X).

Everything else (iNTTs, lookup permutation, grand-product scans, openings, the IPA) is replicated: it needs the whole column
set on every rank, which this backend has by construction (k <= 20 fits one B200; the streamed exchange for k = 22 is
sharded_model.py's).  The proof bytes are identical on every rank and identical to the one-GPU proof.

ShardedCommits is a mixin over any backend of plonk.create_proof, so the commit partition is tested on CPU over gloo with the
oracle's PythonBackend (tests/test_parallel_cpu.py); ShardedGpuBackend adds the coset partition of plonk.GpuBackend."""
from __future__ import annotations

import numpy as np

from . import parallel
from .plonk import GpuBackend


class ShardedCommits:
    """mixin: commit_lagrange_many / commit_many sharded by column.  Needs self.dist (torch.distributed or None) and
    self.comm_device ("cuda" for nccl, "cpu" for gloo)."""
    dist = None
    comm_device = "cpu"

    def _sharded_points(self, commit_many, vecs, blinds):
        d = self.dist
        if d is None or d.get_world_size() == 1 or not len(vecs):
            return commit_many(vecs, blinds)
        import torch
        world, rank = d.get_world_size(), d.get_rank()
        mine = parallel.shard_columns(len(vecs), world, rank)
        local = commit_many([vecs[i] for i in mine], [blinds[i] for i in mine]) if mine else []
        per_rank = (len(vecs) + world - 1) // world
        buf = np.zeros((per_rank, 65), dtype=np.uint8)               # flag, x, y (32-byte little-endian); flag 0 = identity
        for j, pt in enumerate(local):
            if pt is not None:
                buf[j, 0] = 1
                buf[j, 1:33] = np.frombuffer(pt[0].to_bytes(32, "little"), dtype=np.uint8)
                buf[j, 33:65] = np.frombuffer(pt[1].to_bytes(32, "little"), dtype=np.uint8)
        t = torch.from_numpy(buf).to(self.comm_device)
        parts = [torch.empty_like(t) for _ in range(world)]
        d.all_gather(parts, t)
        parts = [p.cpu().numpy() for p in parts]
        out = []
        for c in range(len(vecs)):
            r, j = parallel.owner_of_column(c, world)
            row = parts[r][j]
            out.append((int.from_bytes(row[1:33].tobytes(), "little"), int.from_bytes(row[33:65].tobytes(), "little")) if row[0] else None)
        return out

    def commit_lagrange_many(self, vecs, blinds):
        return self._sharded_points(super().commit_lagrange_many, vecs, blinds)

    def commit_many(self, vecs, blinds):
        return self._sharded_points(super().commit_many, vecs, blinds)


class ShardedGpuBackend(ShardedCommits, GpuBackend):
    """plonk.GpuBackend on this process's GPU + the partitions above over `dist` (an initialised NCCL process group)"""
    comm_device = "cuda"

    def __init__(self, ctx, k, cs_degree, dist=None, params=None):
        super().__init__(ctx, k, cs_degree, params=params)
        self.dist = dist

    def _my_cosets(self, ncos):
        d = self.dist
        if d is None or d.get_world_size() == 1:
            return list(range(ncos))
        return parallel.shard_cosets(ncos, d.get_world_size(), d.get_rank())

    def _exchange_cosets(self, vals, mine):
        d = self.dist
        if d is None or d.get_world_size() == 1:
            return vals
        t = self.torch
        local = vals[mine] if mine else t.zeros((0,) + tuple(vals.shape[1:]), dtype=vals.dtype, device=vals.device)
        self._sync()                      # the library's stream wrote vals; NCCL runs on torch's
        out = parallel.all_gather_columns(local.contiguous(), vals.shape[0], d).contiguous()
        self._sync()
        return out
