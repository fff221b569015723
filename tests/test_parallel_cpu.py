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


# ---- coset split of the quotient evaluation ---------------------------------------------------------------
def _coset_worker(rank, world, port, n_cols, q):
    import random
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.load_package()
    from tiny_ram_halo2_b200 import parallel as PL, poly as P
    from util import pm
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        F = pm.Fp
        dom = pm.EvaluationDomain(F, 4, 3)                     # n = 8, 4 cosets of size 8 in the extended domain
        n, ncos = dom.n, 1 << (dom.extended_k - dom.k)
        rnd = random.Random(5)
        coeff = [[rnd.randrange(F.p) for _ in range(n)] for _ in range(n_cols)]
        leaves = [P.Poly(i % n_cols) for i in range(4)]
        ast = leaves[0] * leaves[1].with_rotation(1) * leaves[2] - leaves[3] * 7 + P.LinearTerm(3) * leaves[0].with_rotation(-1)
        to_t = lambda cols: (torch.from_numpy(np.stack([O.ints_to_limbs(c) for c in cols]).view(np.int64)) if len(cols)
                             else torch.zeros((0, n, 4), dtype=torch.int64))
        to_i = lambda t: [O.limbs_to_ints(c.numpy().view(np.uint64)) for c in t]

        def eval_coset(all_coeff, cs):
            ext = [dom.coeff_to_extended(c) for c in to_i(all_coeff)]
            return to_t([pm.evaluate_ast(dom, ast, ext)[cs::ncos]])[0]

        def combine(vals):
            v = to_i(vals)
            h_ext = [v[i % ncos][i // ncos] for i in range(dom.extended_len())]
            return dom.extended_to_coeff(h_ext)

        mine = PL.shard_columns(n_cols, world, rank)
        got = PL.quotient_cosets_sharded(to_t([coeff[c] for c in mine]), n_cols, ncos, eval_coset, combine, dist)
        want = dom.extended_to_coeff(pm.evaluate_ast(dom, ast, [dom.coeff_to_extended(c) for c in coeff]))
        # the gather must also restore global column order
        order_ok = to_i(PL.all_gather_columns(to_t([coeff[c] for c in mine]), n_cols, dist)) == coeff
        # the streamed (blockwise) exchange delivers the same columns in the same global order
        seen = []
        for g0, blk in PL.all_gather_column_blocks(to_t([coeff[c] for c in mine]), n_cols, 2, dist):
            order_ok = order_ok and g0 == len(seen)
            seen += to_i(blk)
        order_ok = order_ok and seen == coeff
        q.put((rank, got == want, order_ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_cols", [5, 4, 1])
def test_coset_split_world2_gloo(n_cols):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_coset_worker, args=(r, 2, port, n_cols, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] and r[2] for r in res), res


def _proof_worker(rank, world, port, q):
    """plonk.create_proof with the commitments sharded over two gloo ranks (sharded_backend.ShardedCommits over the oracle's
    PythonBackend): every rank must produce the bytes of the unsharded proof"""
    import random
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.load_package()
    from tiny_ram_halo2_b200 import plonk as PL, sharded_backend as SB
    from util import pm
    import plonk_model as VM
    import plonk_circuits
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        C = pm.Vesta

        class ShardedPython(SB.ShardedCommits, VM.PythonBackend):
            pass

        proofs = []
        for sharded in (True, False):
            cs, fixed, copies, adv, inst = plonk_circuits.standard(PL, with_lookup=True, wide_lookup=True)
            be = ShardedPython(C, 4, cs.degree())
            be.dist = dist if sharded else None
            pk = PL.keygen(be, cs, fixed, copies)
            rnd = random.Random(3)
            proofs.append((PL.create_proof(be, pk, inst, adv, lambda: rnd.randrange(C.scalar.p), PL.Blake2bWrite(C.base.p, C.scalar.p)),
                           pk.vk.fixed_commitments, pk.vk.permutation_commitments))
        ok = proofs[0] == proofs[1] and VM.verify_proof(C, be.params, pk.vk, inst, proofs[0][0])
        q.put((rank, bool(ok), proofs[0][0]))
    finally:
        dist.destroy_process_group()


def test_sharded_commitments_give_the_same_proof_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_proof_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=400) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), [r[:2] for r in res]
    assert res[0][2] == res[1][2]


def _row_exchange_worker(rank, world, port, q):
    """the round-2 partitions of sharded_backend.ShardedGpuBackend over gloo on CPU tensors: block ownership + in-place
    all_gather, the row-slice pack + all-to-all of the quotient, the chunk prefixes of the permutation argument, ShardedRng"""
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.load_package()
    from tiny_ram_halo2_b200 import parallel as PL, sharded_backend as SB
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok = True
        # 1. blocks + in-place all_gather: 7 columns, rank r fills its block only
        n, ncols = 32, 7
        want = torch.arange(ncols * n * 4, dtype=torch.int64).reshape(ncols, n, 4)
        per, lo, hi = PL.block_range(ncols, world, rank)
        buf = torch.full((per * world, n, 4), -1, dtype=torch.int64)
        buf[lo:hi] = want[lo:hi]
        PL.all_gather_blocks_inplace(buf, per, dist)
        ok &= bool(torch.equal(buf[:ncols], want))
        # 1b. the deferred exchange of lagrange_to_coeff_many: whole blocks only, left-over columns filled by everybody, the
        #     all_gather in flight while the NEXT batch writes the slot right behind this one
        per_w, main, lo_w, hi_w = PL.whole_block_range(ncols, world, rank)
        arena = torch.full((ncols + 2, n, 4), -1, dtype=torch.int64)
        arena[lo_w:hi_w] = want[lo_w:hi_w]
        arena[main:ncols] = want[main:ncols]
        work = PL.all_gather_blocks_inplace(arena[:main], per_w, dist, async_op=True)
        arena[ncols] = 12345                                  # the next batch's first slot
        if work is not None:
            work.wait()
        ok &= bool(torch.equal(arena[:ncols], want)) and bool((arena[ncols] == 12345).all()) and bool((arena[ncols + 1] == -1).all())
        ok &= per_w * world == main <= ncols < main + world and (work is not None) == (per_w > 0)
        # 2. the quotient's exchange: every rank ends up with rows [row0 - H, row0 + S + H) (cyclic) of EVERY column
        H = 3
        row0, S = PL.row_slice_bounds(n, world, rank)
        own = torch.zeros((max(per, 1), n, 4), dtype=torch.int64)
        own[:hi - lo] = want[lo:hi]
        send = torch.full((world, max(per, 1), S + 2 * H, 4), -7, dtype=torch.int64)
        recv = torch.empty_like(send)
        PL.pack_row_slices(own, hi - lo, n, world, H, H, send)
        PL.exchange_row_slices(send, recv, dist)
        flat = recv.view(world * max(per, 1), S + 2 * H, 4)
        rows = (torch.arange(S + 2 * H) + row0 - H) % n
        for c in range(ncols):
            ok &= bool(torch.equal(flat[c], want[c][rows]))
        # 3. ShardedRng: one stream on every rank, seeded by rank 0
        p = 0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001
        rng = SB.ShardedRng(p, dist, "cpu")
        mine = ([rng() for _ in range(5)], rng.vector(4).tolist(), rng.seed)
        every = [None] * world
        dist.all_gather_object(every, mine)
        ok &= all(e == every[0] for e in every) and all(0 <= v < p for v in mine[0]) and len(set(mine[0])) == 5
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_round2_partitions_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_row_exchange_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res


def test_chunk_prefixes_reproduce_the_chained_products():
    """chunk i computed from 1 and scaled by the product of the earlier chunks' ends == the chain z_i[0] = z_{i-1}[u]"""
    import random
    import __graft_entry__ as ge
    ge.load_package()
    from tiny_ram_halo2_b200 import parallel as PL
    p = 0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001
    rnd = random.Random(9)
    u, chunks = 10, 4
    ratios = [[rnd.randrange(1, p) for _ in range(u)] for _ in range(chunks)]
    chained, start = [], 1
    for r in ratios:
        z = [start]
        for x in r:
            z.append(z[-1] * x % p)
        chained.append(z); start = z[u]
    local = []
    for r in ratios:
        z = [1]
        for x in r:
            z.append(z[-1] * x % p)
        local.append(z)
    pref = PL.chunk_prefixes([z[u] for z in local], p)
    assert [[pref[i] * v % p for v in local[i]] for i in range(chunks)] == chained
    assert PL.block_range(47, 8, 7) == (6, 42, 47) and PL.block_range(5, 8, 6) == (1, 5, 5) and PL.block_range(0, 2, 1) == (0, 0, 0)


def test_sharded_rng_is_a_seeded_stream():
    import __graft_entry__ as ge
    ge.load_package()
    from tiny_ram_halo2_b200 import sharded_backend as SB
    p = 0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001
    a, b, c = SB.ShardedRng(p, seed=b"\x01" * 32), SB.ShardedRng(p, seed=b"\x01" * 32), SB.ShardedRng(p, seed=b"\x02" * 32)
    xs = [a() for _ in range(4)]
    assert xs == [b() for _ in range(4)] != [c() for _ in range(4)]
    v = a.vector(8)
    assert v.shape == (8, 4) and (v == b.vector(8)).all() and int(v[:, 3].max()) < 1 << 62
    assert len(SB.ShardedRng(p).seed) == 32           # OS seed by default
