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
