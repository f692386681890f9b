"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: trajectory / frequency sharding,
the single all-reduce of heat-current sums, the all-gather of T(w) blocks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sclmd_b200 import parallel as PAR


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)                       # same data on every rank, each takes its shard
        ntraj, nmd, nb = 7, 16, 2
        cur = rng.standard_normal((nb, ntraj, nmd))
        lo, hi = PAR.shard_range(ntraj, rank, world)
        sums = [cur[b, lo:hi].sum() for b in range(nb)]
        means = PAR.ensemble_mean_currents(sums, (hi - lo) * nmd)
        want = cur.reshape(nb, -1).mean(axis=1) * 243414.0
        ok1 = np.allclose(means, want, rtol=1e-13, atol=0)
        # frequency blocks
        nw = 11
        tm = rng.uniform(0, 3, nw)
        wlo, whi = PAR.shard_range(nw, rank, world)
        full = PAR.gather_blocks(tm[wlo:whi], nw)
        ok2 = np.array_equal(full, tm)
        h = 0.37
        part = PAR.trapezoid_partial(tm[wlo:whi], wlo, nw, h)
        tot = PAR.allreduce_sum([part])[0]
        ok3 = abs(tot - h / 2 * (2 * tm.sum() - tm[0] - tm[-1])) < 1e-13
        q.put((rank, ok1, ok2, ok3))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] and r[3] for r in res), res


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 1024, 100001):
        for world in (1, 2, 3, 8):
            blocks = [PAR.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_is_identity():
    assert np.array_equal(PAR.allreduce_sum([1.0, 2.0]), [1.0, 2.0])
    assert np.array_equal(PAR.gather_blocks([1.0, 2.0, 3.0], 3), [1.0, 2.0, 3.0])
    m = PAR.ensemble_mean_currents([2.0, -2.0], 4, curcof=1.0)
    assert np.allclose(m, [0.5, -0.5])
    assert PAR.thermal_conductance(m, 300.0, 0.1) == pytest.approx(0.5 / 30.0)


# ---------------------------------------------------------------- the md / bpt classes inside a distributed job (host logic)
def _facade_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import contextlib
        import io
        from sclmd_b200.md import md
        from sclmd_b200.synthetic import spring_chain_dyn
        natoms, ntraj = 6, 7
        K = spring_chain_dyn(natoms, seed=3)
        np.random.seed(100 if rank == 0 else 999)            # only rank 0's stream may matter
        with contextlib.redirect_stdout(io.StringIO()):
            m = md(0.3, 16, 300.0, axyz=[["C", float(i), 0.0, 0.0] for i in range(natoms)], dyn=K, ntraj=ntraj)
            m.initialise()
        seed = PAR.broadcast_int(np.random.randint(0, 2 ** 62))
        # a frequency sweep through the sharding helper (what bpt.gettm / sig.getse do)
        om = np.linspace(0.0, 2.0, 11)
        full = PAR.sharded_sweep(lambda w: np.stack([np.cos(w), np.sin(w)], axis=1) * (1 + 1j), om)
        q.put((rank, m.traj0, m.ntraj, m.ntraj_global, m.sharded, np.array(m.q), np.array(m.p), seed, full))
    finally:
        dist.destroy_process_group()


def test_md_class_shards_its_ensemble_inside_a_distributed_job():
    """md(..., ntraj=N) under torch.distributed (gloo here, NCCL on the GPUs): contiguous trajectory blocks, initial conditions cut
    from rank 0's draw for the whole ensemble, one broadcast seed; the sweep helper returns the full grid on every rank"""
    import contextlib
    import io
    from sclmd_b200.md import md
    from sclmd_b200.synthetic import spring_chain_dyn
    world, natoms, ntraj = 2, 6, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_facade_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda r: r[0])
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        for p in procs:
            if p.is_alive():
                p.kill()
    # the same ensemble in ONE process with rank 0's random stream
    np.random.seed(100)
    with contextlib.redirect_stdout(io.StringIO()):
        one = md(0.3, 16, 300.0, axyz=[["C", float(i), 0.0, 0.0] for i in range(natoms)], dyn=spring_chain_dyn(natoms, seed=3), ntraj=ntraj)
        one.initialise()
    assert not one.sharded and one.traj0 == 0 and one.ntraj == ntraj
    seed_one = int(np.random.randint(0, 2 ** 62))
    blocks = [(r[1], r[2]) for r in res]
    assert blocks == [(0, 4), (4, 3)] and all(r[3] == ntraj and r[4] for r in res)
    assert np.array_equal(np.concatenate([r[5] for r in res]), one.q) and np.array_equal(np.concatenate([r[6] for r in res]), one.p)
    assert res[0][7] == res[1][7] == seed_one
    om = np.linspace(0.0, 2.0, 11)
    want = np.stack([np.cos(om), np.sin(om)], axis=1) * (1 + 1j)
    assert all(np.array_equal(r[8], want) for r in res)


def _stream_k_schedule(ntiles, KI, grid):
    """host-side restatement of the work split of dgemm_tma_kernel (sclmd_b200/csrc/dgemm_tma.cuh): equal contiguous ranges of the
    (tile, K-slab) space, each CTA walking its range from the last tile to the first"""
    total = ntiles * KI
    ctas = []
    for bid in range(grid):
        beg, end = total * bid // grid, total * (bid + 1) // grid
        pieces, hi = [], end
        while hi > beg:
            tile = (hi - 1) // KI
            tstart = tile * KI
            lo = max(beg, tstart)
            pieces.append((tile, lo - tstart, hi - tstart))
            hi = lo
        ctas.append(pieces)
    return ctas


@pytest.mark.parametrize("ntiles,KI,grid", [(152, 38, 148), (80, 188, 148), (376, 38, 148), (1, 5, 5), (7, 3, 4), (65552, 19, 148), (3, 100, 148)])
def test_stream_k_schedule_covers_every_tile_once_and_waits_only_on_earlier_ctas(ntiles, KI, grid):
    """the properties the fix-up of the TMA GEMM relies on: the pieces of a tile tile it exactly; one CTA -- the one holding the tile's
    end -- finalises it and needs partials only from CTAs with a LOWER logical index (tickets are drawn at start: those CTAs are already
    running, so a plain launch cannot deadlock); a CTA stores at most one partial, and stores it before any piece it has to wait for"""
    grid = min(grid, ntiles * KI)
    ctas = _stream_k_schedule(ntiles, KI, grid)
    cover = {}
    for bid, pieces in enumerate(ctas):
        assert pieces, bid
        partial_steps = [i for i, (t, kb, ke) in enumerate(pieces) if ke < KI]
        assert len(partial_steps) <= 1 and all(i == 0 for i in partial_steps)      # the partial is the FIRST thing a CTA does
        for tile, kb, ke in pieces:
            cover.setdefault(tile, []).append((kb, ke, bid))
    assert sorted(cover) == list(range(ntiles))
    for tile, segs in cover.items():
        segs.sort()
        assert segs[0][0] == 0 and segs[-1][1] == KI
        for (a0, a1, _), (b0, b1, _) in zip(segs, segs[1:]):
            assert a1 == b0                                                         # no gap, no overlap
        finaliser = segs[-1][2]
        assert all(bid < finaliser for _, _, bid in segs[:-1])                      # waits only on CTAs that drew an earlier ticket
