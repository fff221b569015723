"""world_size-2 gloo tests of the multi-GPU sharding logic (tiny-ram-halo2_b200/parallel.py) on CPU: the oracle stands in
for the device so that only the partitioning / gather / ordering logic is under test."""
import os
import socket

import numpy as np
import pytest

from util import O, make_points, scalars_uniform, affine_of


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _jac(curve, aff):
    """affine (8,) -> normalised Jacobian (3, 4) as the C ABI returns it"""
    out = np.zeros((3, 4), dtype=np.uint64)
    if aff.any():
        out[0], out[1] = aff[:4], aff[4:]
        out[2] = O.to_mont(O.BASE_FIELD[curve], O.ints_to_limbs([1]))[0]
    return out


def _worker(rank, world, port, n_cols, n, q):
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.load_package()
    from tiny_ram_halo2_b200 import parallel as PL
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        curve = O.VESTA
        pts = make_points(curve, n)
        cols = scalars_uniform(curve, n_cols * n, 7).reshape(n_cols, n, 4)
        commit = lambda cs: np.stack([_jac(curve, O.msm(curve, c, pts)) for c in cs])
        mine = PL.shard_columns(n_cols, world, rank)
        got = PL.commit_columns_sharded(cols[mine], n_cols, commit, dist)
        want = commit(cols)
        ok_cols = bool(np.array_equal(got, want))
        # point-range split of ONE msm
        lo, hi = PL.split_point_range(n, world, rank)
        def points_sum(parts):
            acc = np.zeros(8, dtype=np.uint64)
            for p in parts:
                acc = O.point_add(curve, acc, affine_of(curve, p))
            return _jac(curve, acc)
        full = PL.msm_point_split(cols[0][lo:hi], lambda s: _jac(curve, O.msm(curve, s, pts[lo:hi])), points_sum, dist)
        ok_split = bool(np.array_equal(full, want[0]))
        q.put((rank, ok_cols, ok_split))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_cols,n", [(5, 257), (2, 64), (1, 33)])
def test_sharding_world2_gloo(n_cols, n):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_cols, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res), res


def test_partitions_cover_exactly():
    import __graft_entry__ as ge
    ge.load_package()
    from tiny_ram_halo2_b200 import parallel as PL
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 7, 8, 1 << 20, (1 << 20) + 1):
            ranges = [PL.split_point_range(n, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in ranges) - min(h - l for l, h in ranges) <= 1
        for n_cols in (0, 1, 5, 497):
            seen = sorted(c for r in range(world) for c in PL.shard_columns(n_cols, world, r))
            assert seen == list(range(n_cols))
            for c in range(n_cols):
                r, j = PL.owner_of_column(c, world)
                assert PL.shard_columns(n_cols, world, r)[j] == c
