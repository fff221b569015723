"""Multi-GPU sharding of the prover hot path: one process per GPU, torch.distributed for the plumbing.

Two natural partitions, neither needs a data-path collective beyond a small gather:
  1. column sharding   -- the ~500 per-proof columns are independent (NTT + commit): round-robin columns over ranks,
                          all_gather the 96-byte commitments.
  2. point-range split -- one large MSM: rank r takes points [r*n/G, (r+1)*n/G) with the matching slice of the (static,
                          pre-sharded) bases; the G partial sums are all_gathered and added (trp_points_sum), exactly how
                          best_multiexp combines its per-thread partial results.
  3. coset split       -- the extended domain is 2^(extended_k - k) cosets of size n and deg h < (j - 1) n, so j - 1 cosets
                          determine the quotient: after ONE all_gather of the coefficient-form columns (the path's only bulk
                          exchange, n_cols * n * 32 B in total), rank r evaluates cosets r, r + G, ... (coset NTT of every
                          column + the quotient program), the n-value results are all_gathered (j - 1 columns in total) and
                          every rank recovers h(X) (trp_dev_cosets_to_coeff) for the column-sharded commits of its pieces.
The functions take the commit / add operations as callables so the same logic runs over NCCL with the CUDA library and
over gloo on CPU in the tests (with the oracle standing in for the device)."""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np


def shard_columns(n_cols: int, world: int, rank: int) -> List[int]:
    """round-robin: column c belongs to rank c % world"""
    return list(range(rank, n_cols, world))


def owner_of_column(col: int, world: int) -> Tuple[int, int]:
    """(rank, local index) of a column under shard_columns"""
    return col % world, col // world


def split_point_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of rank's slice; slices differ by at most one point and cover [0, n) in rank order"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _all_gather(t, dist):
    import torch
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return out


def commit_columns_sharded(columns_local, n_cols: int, commit: Callable, dist=None, device="cpu"):
    """columns_local: this rank's columns (in shard_columns order).  Returns the (n_cols, 3, 4) uint64 commitments of ALL
    columns in global order on every rank.  commit(cols) -> (len(cols), 3, 4) uint64."""
    import torch
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    mine = shard_columns(n_cols, world, rank)
    per_rank = (n_cols + world - 1) // world
    local = np.zeros((per_rank, 3, 4), dtype=np.uint64)
    if len(mine):
        local[:len(mine)] = np.asarray(commit(columns_local), dtype=np.uint64).reshape(len(mine), 3, 4)
    if world == 1:
        return local[:n_cols]
    t = torch.from_numpy(local.view(np.int64)).to(device)
    parts = [p.cpu().numpy().view(np.uint64) for p in _all_gather(t, dist)]
    out = np.zeros((n_cols, 3, 4), dtype=np.uint64)
    for c in range(n_cols):
        r, j = owner_of_column(c, world)
        out[c] = parts[r][j]
    return out


def msm_point_split(scalars_local, msm_local: Callable, points_sum: Callable, dist=None, device="cpu"):
    """scalars_local: this rank's slice of the scalars (split_point_range).  msm_local(scalars) -> (3, 4) partial sum over the
    rank's slice of the bases; points_sum((G, 3, 4)) -> (3, 4).  Every rank returns the full MSM."""
    import torch
    part = np.asarray(msm_local(scalars_local), dtype=np.uint64).reshape(3, 4)
    if dist is None or dist.get_world_size() == 1:
        return points_sum(part.reshape(1, 3, 4))
    t = torch.from_numpy(part.view(np.int64).copy()).to(device)
    parts = np.stack([p.cpu().numpy().view(np.uint64) for p in _all_gather(t, dist)])
    return points_sum(parts)


def shard_cosets(n_cosets: int, world: int, rank: int) -> List[int]:
    """round-robin: coset j belongs to rank j % world"""
    return list(range(rank, n_cosets, world))


def all_gather_columns(local, n_cols: int, dist=None):
    """local: torch tensor (len(shard_columns(n_cols, world, rank)), ...) of this rank's columns, on the device the process
    group communicates from (cuda for nccl, cpu for gloo).  Returns (n_cols, ...) in GLOBAL column order on every rank."""
    import torch
    if dist is None or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    per_rank = (n_cols + world - 1) // world
    mine = len(shard_columns(n_cols, world, rank))
    if local.shape[0] != mine:
        raise ValueError(f"rank {rank} holds {local.shape[0]} columns, expected {mine}")
    padded = local
    if mine < per_rank:
        padded = torch.zeros((per_rank,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded[:mine].copy_(local)
    gathered = torch.empty((world, per_rank) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather(list(gathered.unbind(0)), padded.contiguous())
    # (rank, slot) -> global column slot * world + rank
    return gathered.transpose(0, 1).reshape((per_rank * world,) + tuple(local.shape[1:]))[:n_cols]


def quotient_cosets_sharded(coeff_local, n_cols: int, n_cosets: int, eval_coset: Callable, combine: Callable, dist=None):
    """coeff_local: this rank's coefficient-form columns (shard_columns order), torch tensor (mine, n, 4).
    eval_coset(all_coeff (n_cols, n, 4), coset) -> (n, 4) tensor: the quotient numerator on that coset;
    combine(vals (n_cosets, n, 4)) -> the quotient's coefficients.  Every rank returns combine's result."""
    import torch
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    all_coeff = all_gather_columns(coeff_local, n_cols, dist)
    mine = shard_cosets(n_cosets, world, rank)
    vals = [eval_coset(all_coeff, cs) for cs in mine]
    if vals:
        local = torch.stack(vals)
    else:
        local = torch.zeros((0,) + tuple(coeff_local.shape[1:]), dtype=coeff_local.dtype, device=coeff_local.device)
    return combine(all_gather_columns(local, n_cosets, dist))


def all_gather_column_blocks(local, n_cols: int, block_slots: int, dist):
    """Streamed form of all_gather_columns: yields (first global column, block) where block holds the global columns
    [g0, g0 + len(block)) in order, block_slots slots of every rank at a time (one all_gather per block; every rank must
    consume the generator in lockstep).  A rank never holds more than one block of foreign columns."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    per_rank = (n_cols + world - 1) // world
    mine = len(shard_columns(n_cols, world, rank))
    if local.shape[0] < mine:
        raise ValueError(f"rank {rank} holds {local.shape[0]} columns, expected {mine}")
    for s0 in range(0, per_rank, block_slots):
        s1 = min(s0 + block_slots, per_rank)
        blk = torch.zeros((s1 - s0,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        hi = min(s1, mine)
        if hi > s0:
            blk[:hi - s0].copy_(local[s0:hi])
        gathered = torch.empty((world,) + tuple(blk.shape), dtype=local.dtype, device=local.device)
        dist.all_gather(list(gathered.unbind(0)), blk)
        g0, g1 = s0 * world, min(s1 * world, n_cols)
        yield g0, gathered.transpose(0, 1).reshape(((s1 - s0) * world,) + tuple(local.shape[1:]))[:g1 - g0].contiguous()


# ---- round 2: every heavy phase of create_proof divided between the devices (sharded_backend.ShardedGpuBackend) -----------------
def block_range(count: int, world: int, rank: int) -> Tuple[int, int, int]:
    """(per, lo, hi): items are owned in contiguous blocks of per = ceil(count / world); rank owns [lo, hi) (possibly empty)"""
    per = -(-count // world) if count else 0
    lo = min(rank * per, count)
    return per, lo, min(lo + per, count)


def whole_block_range(k: int, world: int, rank: int) -> Tuple[int, int, int, int]:
    """(per, main, lo, hi): columns [lo, hi) = [rank * per, (rank + 1) * per) of a batch of k belong to `rank`, per = k // world; the
    k - main left-over columns (main = per * world) belong to EVERY rank.  An exchange of [0, main) in blocks of per never touches
    anything outside the batch, so it may stay in flight while the slots behind the batch are being written."""
    per = k // world
    return per, per * world, rank * per, (rank + 1) * per


def all_gather_blocks_inplace(buf, per: int, dist, async_op: bool = False):
    """buf: (>= per * world, ...) tensor whose block [rank * per, (rank + 1) * per) this rank has filled; after the call every
    rank holds every block (NCCL / gloo in-place all-gather: the send buffer is the rank's slot of the receive buffer).
    async_op: return the collective's work handle instead (None if nothing was started); the caller must wait() on it before
    anybody reads the other ranks' blocks, and must not write to buf until then."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if per == 0 or world == 1:
        return None if async_op else buf
    out = buf[:per * world]
    work = dist.all_gather_into_tensor(out, out[rank * per:(rank + 1) * per], async_op=async_op)
    return work if async_op else buf


def row_slice_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """rows of a size-n coset a rank evaluates in the row-split quotient: [rank * n / world, (rank + 1) * n / world)"""
    if n % world:
        raise ValueError("the number of devices must divide the domain size")
    s = n // world
    return rank * s, s


def pack_row_slices(own, count: int, n: int, world: int, halo_before: int, halo_after: int, send):
    """own: (>= count, n, ...) full-length columns; send: (world, per, S + halo_before + halo_after, ...).  Fills
    send[d, :count] with rows [d * S - halo_before, (d + 1) * S + halo_after) (cyclic) of the first `count` columns: what
    device d needs of these columns to evaluate its rows with rotations in [-halo_before, +halo_after]."""
    s = n // world
    width = s + halo_before + halo_after
    if halo_before > n or halo_after > n:
        raise ValueError("halo larger than the domain")
    for d in range(world):
        lo = d * s - halo_before
        pos = 0
        while pos < width:                                   # at most three contiguous pieces (wrap at either end)
            src = (lo + pos) % n
            run = min(width - pos, n - src)
            send[d, :count, pos:pos + run].copy_(own[:count, src:src + run])
            pos += run
    return send


def exchange_row_slices(send, recv, dist):
    """all-to-all of the packed slices: recv[src] = the (per, S + halo, ...) block device `src` packed for this rank"""
    if dist is None or dist.get_world_size() == 1:
        recv.copy_(send)
        return recv
    dist.all_to_all_single(recv, send)
    return recv


def chunk_prefixes(ends: Sequence[int], p: int) -> List[int]:
    """permutation grand products computed chunk by chunk, each from 1: chunk i of the chained argument is
    prefix_i * local_i with prefix_i = prod_{j < i} local_j[last usable row] (halo2 starts chunk i at z_{i-1}[u])"""
    out, acc = [], 1
    for e in ends:
        out.append(acc)
        acc = acc * e % p
    return out
