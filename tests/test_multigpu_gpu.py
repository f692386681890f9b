"""The md / bpt classes inside a 2-rank NCCL job (one process per GPU): md(..., ntraj=N) shards its ensemble, bath.gnoi() draws the
GLOBAL trajectory streams from one broadcast seed, mean_currents() is the one all-reduce of the path, bpt.gettm() cuts the frequency
grid into blocks and all-gathers -- results equal the same job on ONE GPU (to rounding: the batched products sum in another order
when the trajectory block changes).  Skipped on boxes with fewer than two GPUs; tests/test_parallel_cpu.py covers the host logic
with gloo."""
import contextlib
import io
import os
import socket

import numpy as np
import pytest

import problems as P

pytestmark = pytest.mark.gpu

NATOMS, NTRAJ, NMD, DT, T0 = 40, 6, 64, 0.25 / 0.658, 300.0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(device):
    from sclmd_b200.md import md
    from sclmd_b200.baths import phbath
    nph = 3 * NATOMS
    K = P.spring_chain_dyn(NATOMS, seed=21)
    m = md(DT, NMD, T0, axyz=[["C", float(i), 0.0, 0.0] for i in range(NATOMS)], dyn=K, ntraj=NTRAJ, device=device)
    gwl, gam = P.gamma_grid(5, 12, 31, wmax=0.3)
    b0 = phbath(T0 * 1.05, list(range(0, 12)), 0.06, 40, DT, NMD, ml=24, gamma=gam, gwl=gwl)       # memory kernel, spectrum on a grid
    b0.gmem()
    b1 = phbath(T0 * 0.95, list(range(nph - 12, nph)), 0.05, 40, DT, NMD)                           # Debye, time-local
    b1.gmem()
    m.AddBath(b0)
    m.AddBath(b1)
    return m


def _run(m):
    np.random.seed(77)
    m.initialise()
    m.ResetHis()
    for b in m.baths:
        b.gnoi()
    m.steps(NMD)
    means, sums, count = m.mean_currents()
    return dict(q=np.array(m.q), p=np.array(m.p), noise=[m.get_noise(i) for i in range(2)], means=means, count=count,
                traj0=m.traj0, ntraj=m.ntraj)


def _bpt(device):
    from sclmd_b200.negf import bpt
    Kn = P.spring_chain_dyn(20, seed=22) / 6.582119569e-4 ** 2
    return bpt(None, 0.25, 0.1, [list(range(6, 18)), list(range(42, 54))], [list(range(0, 6)), list(range(54, 60))], dynmatfile=Kn, num=37,
               device=device)


def _worker(rank, world, port, q, tmp):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["LOCAL_RANK"] = str(rank)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        os.chdir(tmp)
        with contextlib.redirect_stdout(io.StringIO()):
            if rank == 1:
                np.random.seed(12345)                # must not matter: seeds and phases come from rank 0
            m = _build(None)                         # device = LOCAL_RANK
            out = _run(m)
            b = _bpt(None)
            b.gettm()
            out["tm"] = np.array(b.tmnumber)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_two_rank_nccl_job_equals_one_gpu(tmp_path, monkeypatch):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    monkeypatch.chdir(tmp_path)
    with contextlib.redirect_stdout(io.StringIO()):
        one = _run(_build(0))
        b = _bpt(0)
        b.gettm()
    tm_one = np.array(b.tmnumber)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = dict(q.get(timeout=150) for _ in range(2))
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        for p in procs:                  # a rank that died leaves the other one inside a collective: do not wait for NCCL's timeout
            if p.is_alive():
                p.kill()
    assert (res[0]["traj0"], res[0]["ntraj"], res[1]["traj0"], res[1]["ntraj"]) == (0, 3, 3, 3)
    rel = lambda a, b: float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))
    for i in range(2):
        both = np.concatenate([res[0]["noise"][i], res[1]["noise"][i]])
        assert rel(both, one["noise"][i]) < 1e-12, i                     # global trajectory streams, one broadcast seed
    assert rel(np.concatenate([res[0]["q"], res[1]["q"]]), one["q"]) < 1e-9
    assert rel(np.concatenate([res[0]["p"], res[1]["p"]]), one["p"]) < 1e-9
    for r in (0, 1):                                                     # the all-reduced ensemble means, on every rank
        assert res[r]["count"] == one["count"] == NTRAJ * NMD
        assert rel(res[r]["means"], one["means"]) < 1e-8
        assert rel(res[r]["tm"], tm_one) < 1e-10                         # frequency blocks, all-gathered
    assert os.path.exists(tmp_path / "transmission.dat")
