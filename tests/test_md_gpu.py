"""Parity of the CUDA ensemble integrator (through the C ABI) against the reference's
golden trajectories and the CPU oracle.  Tolerance (north_star): <= 1e-10 relative per
step on q,p given identical injected noise; <= 1e-8 on ensemble observables."""
import os

import numpy as np
import pytest

import problems as P
from oracle import sclmd_oracle as O

pytestmark = pytest.mark.gpu

TOL_STEP = 1e-10
TOL_OBS = 1e-8


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def bath_args(c, b):
    """(kernel, Mq, Mp) exactly as sclmd_b200.baths hands them to the engine."""
    if c["kinds"][b] == "ph":
        return c["kern"][b], None, None
    e = c["e"]
    sym = lambda a: None if a is None else 0.5 * (a + a.T)
    asym = lambda a: None if a is None else 0.5 * (a - a.T)
    nc = len(c["cids"][b])
    z = np.zeros((nc, nc))
    exim = z if e["exim"][b] is None else asym(e["exim"][b])
    zeta1 = z if e["zeta1"][b] is None else sym(e["zeta1"][b])
    zeta2 = z if e["zeta2"][b] is None else asym(e["zeta2"][b])
    kern = np.array([sym(e["efric"][b])])
    if exim.any() and zeta1.any() and zeta2.any():        # baths.py:233
        return kern, e["bias"][b] * (exim - zeta1), -e["bias"][b] * zeta2
    return kern, None, None


@pytest.mark.parametrize("name", list(P.MD_CASES))
def test_golden_trajectories(name, golden_dir):
    from sclmd_b200.engine import MDEngine
    c = P.MD_CASES[name]()
    g = np.load(os.path.join(golden_dir, "md_%s.npz" % name))
    K = P.psd_project(c["K"])
    ntraj = 3
    eng = MDEngine(K.shape[0], ntraj, c["dt"], c["nmd"])
    eng.set_dyn(K)
    if c["cons"] is not None:
        eng.set_constraint([i for grp in c["cons"] for i in grp])
    for b in range(len(c["cids"])):
        kern, Mq, Mp = bath_args(c, b)
        eng.add_bath(c["cids"][b], kern, Mq, Mp)
        eng.set_noise(b, np.stack([0.5 * c["noise"][b], c["noise"][b], -c["noise"][b]]))
    eng.set_state(g["q0"], g["p0"], 0)
    n = int(g["nsteps"])
    full = g["q"].shape[0] == n
    for s in range(n):
        eng.run(1)
        if full or s == n - 1:
            q, p, t = eng.get_state()
            k = s if full else 0
            assert t == s + 1
            assert relerr(q[1], g["q"][k]) < TOL_STEP, (name, s)
            assert relerr(p[1], g["p"][k]) < TOL_STEP, (name, s)
    for b in range(len(c["cids"])):
        assert relerr(eng.current(b)[1], g["cur"][b]) < TOL_OBS
        assert relerr(eng.current_sums(b)[1], g["cur"][b].sum()) < TOL_OBS
    assert relerr(eng.etot()[1], g["etot"]) < TOL_STEP
    eng.close()


@pytest.mark.parametrize("kind,ml,nc,ntraj,nsteps", [
    ("diag", 300, 30, 7, 330),      # ring wraps, T=4 tiles with a remainder, several splits
    ("diag", 2, 6, 1, 9),           # shortest memory, single trajectory
    ("diag", 1500, 150, 9, 40),     # C1-sized bath, long memory
    ("full", 40, 30, 5, 90),        # split-K full-kernel tail
    ("full", 2, 7, 2, 11),          # odd nc -> padded rows
    ("full", 6, 30, 70, 12),        # > 64 trajectories, narrow output: 64-wide GEMM tiles with a wave-fitting split count
    ("full", 3, 150, 130, 8),       # >= 96 trajectories: TMA / stream-K contraction over the ring, 160-wide tiles
    ("full", 37, 34, 100, 80),      # the same kernel with a ring that wraps (80 steps > ml), ragged 16-wide K slabs (ncp = 34)
])
def test_memory_kernels_vs_oracle(kind, ml, nc, ntraj, nsteps):
    from sclmd_b200.engine import MDEngine
    natoms = max(12, (2 * nc + 8) // 3 + 2)
    nph = 3 * natoms
    dt, nmd = 0.25 / 0.658, 64
    K = P.psd_project(P.spring_chain_dyn(natoms, seed=5))
    cons = [list(range(0, 3)), list(range(nph - 3, nph))]
    cids = [list(range(3, 3 + nc)), list(range(nph - 3 - nc, nph - 3))]
    eng = MDEngine(nph, ntraj, dt, nmd)
    ens = O.EnsembleMD(K, dt, nmd, ntraj, cons)
    eng.set_dyn(K)
    eng.set_constraint([i for grp in cons for i in grp])
    for b in range(2):
        kern = P.diag_kernel(ml, nc, dt, 50 + b) if kind == "diag" else P.full_kernel(ml, nc, dt, 50 + b)
        nz = P.injected_noise(ntraj, nmd, nc, seed=60 + b)
        eng.add_bath(cids[b], kern)
        eng.set_noise(b, nz)
        ens.add_bath(cids[b], kern, nz)
    rng = np.random.default_rng(7)
    q0, p0 = 0.05 * rng.standard_normal((ntraj, nph)), 0.02 * rng.standard_normal((ntraj, nph))
    eng.set_state(q0, p0, 0)
    ens.q[:], ens.p[:] = q0, p0
    done = 0
    for chunk in (1, 2, nsteps - 3):
        eng.run(chunk)
        ens.run(chunk)
        done += chunk
        q, p, t = eng.get_state()
        assert t == done
        assert relerr(q, ens.q) < TOL_STEP and relerr(p, ens.p) < TOL_STEP, (kind, ml, done)
    for b in range(2):
        assert relerr(eng.current(b), ens.baths[b]["cur"]) < TOL_OBS
        # history read-back in the reference's order (row 0 newest): bit-exact indexing
        h = eng.get_history(b)
        want = np.stack([ens.baths[b]["ring"][:, (done - 1 - i) % ml, :] for i in range(ml)], axis=1)
        assert relerr(h, want) < TOL_STEP
    assert relerr(eng.etot(), ens.etot) < TOL_STEP
    eng.close()


@pytest.mark.parametrize("ml,nc,ntraj", [(200, 22, 5), (4096, 12, 4), (131, 7, 1)])
def test_time_blocked_tails_equal_direct_tails(ml, nc, ntraj):
    """the time-blocked far/near split of the friction tail is the same sum in another order: the tensor-pipe far pass over 32-step
    blocks (k_tail_far_mma, needs ml % 8 == 0: the ml = 131 case falls back to the 16-step kernel), the direct pass, the 32-step DFMA
    kernel and the 16-step DFMA kernel against the oracle, with mode switches and run() calls that end mid-block"""
    from sclmd_b200.engine import MDEngine
    natoms = 12
    nph, dt, nmd = 3 * natoms, 0.25 / 0.658, 64
    K = P.psd_project(P.spring_chain_dyn(natoms, seed=9))
    kern = P.diag_kernel(ml, nc, dt, 4, tau=400.0)          # slowly decaying memory: old history matters
    nz = P.injected_noise(ntraj, nmd, nc, seed=5)
    rng = np.random.default_rng(1)
    q0, p0 = 0.05 * rng.standard_normal((ntraj, nph)), 0.02 * rng.standard_normal((ntraj, nph))
    ens = O.EnsembleMD(K, dt, nmd, ntraj, None)
    ens.add_bath(list(range(nc)), kern, nz)
    ens.q[:], ens.p[:] = q0, p0
    engs = []
    for mode in (1, 0, 4, 5):           # tensor-pipe 32-step blocks, direct, 32-step blocks (k_tail_far_wsx), 16-step blocks (k_tail_far_ws)
        e = MDEngine(nph, ntraj, dt, nmd)
        e.set_dyn(K)
        e.add_bath(list(range(nc)), kern)
        e.set_noise(0, nz)
        e.set_tail_block(mode)
        e.set_state(q0, p0, 0)
        engs.append(e)
    done = 0
    for chunk in (5, 16, 1, 27, 40):
        ens.run(chunk)
        done += chunk
        for e in engs:
            e.run(chunk)
            q, p, t = e.get_state()
            assert t == done and relerr(q, ens.q) < TOL_STEP and relerr(p, ens.p) < TOL_STEP, (ml, done)
    engs[0].set_tail_block(0)        # switch modes in the middle of a block
    engs[1].set_tail_block(4)
    engs[2].set_tail_block(1)
    engs[3].set_tail_block(1)
    ens.run(23)
    for e in engs:
        e.run(23)
        q, p, _ = e.get_state()
        assert relerr(q, ens.q) < TOL_STEP and relerr(p, ens.p) < TOL_STEP
    # restart into a fresh blocked engine at a time that is not a block boundary
    q, p, t = engs[0].get_state()
    f = MDEngine(nph, ntraj, dt, nmd)
    f.set_dyn(K)
    f.add_bath(list(range(nc)), kern)
    f.set_noise(0, nz)
    f.set_state(q, p, t)
    f.set_history(0, engs[0].get_history(0))
    ens.run(19)
    f.run(19)
    qf, pf, _ = f.get_state()
    assert relerr(qf, ens.q) < TOL_STEP and relerr(pf, ens.p) < TOL_STEP
    for e in engs + [f]:
        e.close()


def test_streamed_noise_rows_and_async_steps_match_resident_table():
    """the end-to-end path of bench.py: per step, noise row t+1 uploaded from the host (asynchronously), one
    asynchronous step, observables read back -- identical to running on a resident noise table"""
    from sclmd_b200.engine import MDEngine
    nph, nc, ml, ntraj, dt, nmd = 36, 8, 150, 5, 0.3, 16
    K = P.psd_project(P.spring_chain_dyn(12, seed=3))
    kern = P.diag_kernel(ml, nc, dt, 1)
    nz = P.injected_noise(ntraj, nmd, nc, seed=2)

    def mk():
        e = MDEngine(nph, ntraj, dt, nmd)
        e.set_dyn(K)
        e.set_constraint([0, 1, 2])
        e.add_bath(list(range(3, 3 + nc)), kern)
        e.set_state(np.full((ntraj, nph), 0.01), np.zeros((ntraj, nph)), 0)
        return e
    a, b = mk(), mk()
    a.set_noise(0, nz)
    rows = np.ascontiguousarray(nz.transpose(1, 0, 2))           # [nmd, ntraj, nc] time-major
    b.set_noise_rows(0, 0, rows[0:1])
    nsteps = 40                                                  # > nmd: the table wraps
    obs_b = []
    for t in range(nsteps):
        b.set_noise_rows(0, (t + 1) % nmd, rows[(t + 1) % nmd:(t + 1) % nmd + 1])
        b.run_async(1)
        obs_b.append(b.step_observables(t % nmd).copy())
    a.run(nsteps)
    qa, pa, ta = a.get_state()
    qb, pb, tb = b.get_state()
    assert ta == tb == nsteps and np.array_equal(qa, qb) and np.array_equal(pa, pb)
    et, cur = a.etot(), a.current(0)
    for t in range(nsteps - nmd, nsteps):
        assert np.array_equal(obs_b[t][0], et[:, t % nmd]) and np.array_equal(obs_b[t][1], cur[:, t % nmd])
    a.close()
    b.close()


def test_lazy_evaluations_bc_equal_the_eager_order(monkeypatch):
    """without constraints the evaluations B, C of a step stay pending and run fused with evaluation A of the next step
    (k_phase_bca), across sclmd_md_run calls too; every access to the state runs them first.  One call, step-by-step asynchronous
    calls with streamed noise rows, step-by-step calls with a state read-back after each (always flushed) and the unfused build
    (SCLMD_NO_FUSE) give bit-identical states, histories and observables"""
    from sclmd_b200.engine import MDEngine
    nph, nc, ml, ntraj, dt, nmd = 36, 8, 150, 5, 0.3, 16
    K = P.psd_project(P.spring_chain_dyn(12, seed=3))
    kern = P.diag_kernel(ml, nc, dt, 1)
    nz = P.injected_noise(ntraj, nmd, nc, seed=2)
    rows = np.ascontiguousarray(nz.transpose(1, 0, 2))
    nsteps = 37

    def mk(resident=True):
        e = MDEngine(nph, ntraj, dt, nmd)
        e.set_dyn(K)
        e.add_bath(list(range(3, 3 + nc)), kern)
        e.add_bath(list(range(20, 20 + nc)), kern[:1])
        e.set_noise(1, 0.5 * nz)
        if resident:
            e.set_noise(0, nz)
        e.set_state(np.full((ntraj, nph), 0.01), np.zeros((ntraj, nph)), 0)
        return e

    def result(e):
        q, p, t = e.get_state()
        return q, p, t, e.get_history(0), e.etot(), e.current(0), e.current(1)

    a = mk()
    a.run(nsteps)
    want = result(a)
    b = mk(resident=False)                                       # streamed rows, asynchronous single steps
    b.set_noise_rows(0, 0, rows[0:1])
    for t in range(nsteps):
        b.set_noise_rows(0, (t + 1) % nmd, rows[(t + 1) % nmd:(t + 1) % nmd + 1])
        b.run_async(1)
        ob = b.step_observables(t % nmd)
        if t >= nsteps - nmd:
            assert np.array_equal(ob[0], want[4][:, t % nmd]) and np.array_equal(ob[1], want[5][:, t % nmd])
    c = mk()                                                     # flushed after every step
    for t in range(nsteps):
        c.run_async(1)
        c.get_state()
    d = mk()                                                     # uneven pieces, a table rewrite (flush) in between
    d.run_async(5)
    d.set_noise(0, nz)
    d.run(11)
    d.run_async(nsteps - 16)
    monkeypatch.setenv("SCLMD_NO_FUSE", "1")
    e = mk()
    monkeypatch.delenv("SCLMD_NO_FUSE")
    e.run(nsteps)
    for other in (b, c, d, e):
        got = result(other)
        assert got[2] == want[2] == nsteps
        for x, y in zip(got, want):
            assert np.array_equal(x, y)
        other.close()
    a.close()


def test_history_roundtrip_and_restart():
    """state + history saved from one engine and loaded into another continue identically
    (md.py:552-562 restart semantics)."""
    from sclmd_b200.engine import MDEngine
    nph, nc, ml, ntraj, dt, nmd = 36, 9, 17, 4, 0.3, 32
    K = P.psd_project(P.spring_chain_dyn(12, seed=3))
    kern = P.diag_kernel(ml, nc, dt, 1)
    nz = P.injected_noise(ntraj, nmd, nc, seed=2)

    def mk():
        e = MDEngine(nph, ntraj, dt, nmd)
        e.set_dyn(K)
        e.add_bath(list(range(nc)), kern)
        e.set_noise(0, nz)
        return e
    a = mk()
    a.set_state(np.full((ntraj, nph), 0.01), np.zeros((ntraj, nph)), 0)
    a.run(25)
    q, p, t = a.get_state()
    b = mk()
    b.set_state(q, p, t)
    b.set_history(0, a.get_history(0))
    a.run(20)
    b.run(20)
    qa, pa, _ = a.get_state()
    qb, pb, _ = b.get_state()
    assert relerr(qb, qa) < 1e-13 and relerr(pb, pa) < 1e-13
    a.close()
    b.close()


def test_errors_are_reported_not_fatal():
    from sclmd_b200.engine import MDEngine
    from sclmd_b200._lib import SclmdError
    e = MDEngine(12, 2, 0.3, 8)
    with pytest.raises(SclmdError):
        e.run(1)                                   # no dynamical matrix: "no driver, no md"
    with pytest.raises(SclmdError):
        e.add_bath([0, 1, 99], np.zeros((1, 3)))   # dof out of range
    with pytest.raises(SclmdError):
        e.add_bath([0, 0], np.zeros((1, 2)))       # duplicate dof
    e.close()


def test_config5_shape_full_size_vs_oracle():
    """BASELINE configs[4] per-trajectory shape at full size (3000 dofs, 2 baths x 300 dofs, diagonal 4096-step memory kernels) with a
    random pre-existing history, so that the whole memory matters from the first step: both history-tail modes (time-blocked with the
    warp-specialised ring pass, and direct) against the oracle across block boundaries and mid-block run() calls"""
    from sclmd_b200.engine import MDEngine
    natoms, nc, ml, ntraj, nmd = 1000, 300, 4096, 4, 64
    nph, dt = 3 * natoms, 0.25 / 0.658
    K = P.spring_chain_dyn(natoms, seed=5)
    cids = [list(range(0, nc)), list(range(nph - nc, nph))]
    kern = [P.diag_kernel(ml, nc, dt, 30 + b, tau=600.0) for b in range(2)]
    nz = [P.injected_noise(ntraj, nmd, nc, seed=40 + b) for b in range(2)]
    rng = np.random.default_rng(50)
    q0, p0 = 0.05 * rng.standard_normal((ntraj, nph)), 0.02 * rng.standard_normal((ntraj, nph))
    hist = [0.02 * rng.standard_normal((ntraj, ml, nc)) for _ in range(2)]      # phis[i] = p_{t-1-i}[cids]
    ens = O.EnsembleMD(K, dt, nmd, ntraj, None)
    for b in range(2):
        ens.add_bath(cids[b], kern[b], nz[b])
        ens.baths[b]["ring"][:, (-1 - np.arange(ml)) % ml, :] = hist[b]
    ens.q[:], ens.p[:] = q0, p0
    ens.t = -1
    for b in ens.baths:
        b["tail"] = ens._tail(b)             # S(0) = dt sum_j k[j] p_{-j}
    ens.t = 0
    engs = []
    for mode in (1, 0):
        e = MDEngine(nph, ntraj, dt, nmd)
        e.set_dyn(K)
        for b in range(2):
            e.add_bath(cids[b], kern[b])
            e.set_noise(b, nz[b])
        e.set_tail_block(mode)
        e.set_state(q0, p0, 0)
        for b in range(2):
            e.set_history(b, hist[b])
        engs.append(e)
    done = 0
    for chunk in (20, 17):
        ens.run(chunk)
        done += chunk
        for e in engs:
            e.run(chunk)
            q, p, t = e.get_state()
            assert t == done and relerr(q, ens.q) < TOL_STEP and relerr(p, ens.p) < TOL_STEP, done
    for b in range(2):
        cur = engs[0].current(b)
        assert relerr(cur[:, :done], ens.baths[b]["cur"][:, :done]) < 1e-8
    for e in engs:
        e.close()


@pytest.mark.parametrize("case", ["one_atom", "unsorted_cids", "overlapping_baths", "bath_on_every_dof", "no_bath", "nmd_wraps"])
def test_edge_shapes_vs_oracle(case):
    """smallest and ragged inputs: a single atom with a one-dof bath, bath dofs given out of order (baths.py:79-80 takes any index
    list), two baths acting on the same dofs (their forces add, md.py:432-434), a bath covering every dof, no bath at all, and
    more steps than nmd (slot indices wrap, md.py:383,397)"""
    from sclmd_b200.engine import MDEngine
    dt = 0.25 / 0.658
    rng = np.random.default_rng(11)
    if case == "one_atom":
        natoms, nmd, ntraj, baths = 1, 8, 1, [([1], "diag", 1)]
    elif case == "unsorted_cids":
        natoms, nmd, ntraj, baths = 6, 16, 3, [([7, 3, 12, 4, 9], "full", 3), ([15, 14, 16], "diag", 4)]
    elif case == "overlapping_baths":
        natoms, nmd, ntraj, baths = 5, 16, 2, [([3, 4, 5, 6], "diag", 2), ([5, 6, 7], "full", 1)]
    elif case == "bath_on_every_dof":
        natoms, nmd, ntraj, baths = 4, 16, 5, [(list(range(12)), "full", 2)]
    elif case == "no_bath":
        natoms, nmd, ntraj, baths = 4, 8, 2, []
    else:
        natoms, nmd, ntraj, baths = 4, 6, 2, [([0, 1, 2], "diag", 5)]
    nph = 3 * natoms
    K = P.psd_project(P.spring_chain_dyn(max(natoms, 2), seed=2))[:nph, :nph]
    K = 0.5 * (K + K.T)
    eng = MDEngine(nph, ntraj, dt, nmd)
    ens = O.EnsembleMD(K, dt, nmd, ntraj, None)
    eng.set_dyn(K)
    for b, (cids, kind, ml) in enumerate(baths):
        nc = len(cids)
        kern = P.diag_kernel(ml, nc, dt, 70 + b) if kind == "diag" else P.full_kernel(ml, nc, dt, 70 + b)
        nz = P.injected_noise(ntraj, nmd, nc, seed=80 + b)
        eng.add_bath(cids, kern)
        eng.set_noise(b, nz)
        ens.add_bath(cids, kern, nz)
    q0, p0 = 0.05 * rng.standard_normal((ntraj, nph)), 0.02 * rng.standard_normal((ntraj, nph))
    eng.set_state(q0, p0, 0)
    ens.q[:], ens.p[:] = q0, p0
    nsteps = 2 * nmd + 3
    eng.run(nsteps)
    ens.run(nsteps)
    q, p, t = eng.get_state()
    assert t == nsteps and relerr(q, ens.q) < TOL_STEP and relerr(p, ens.p) < TOL_STEP
    assert relerr(eng.etot(), ens.etot) < TOL_STEP
    for b in range(len(baths)):
        assert relerr(eng.current(b), ens.baths[b]["cur"]) < TOL_OBS
    eng.run(0)                                   # zero steps: a no-op
    assert eng.get_state()[2] == nsteps
    eng.close()


@pytest.mark.parametrize("ntraj,cons", [(1, True), (2, True), (2, False), (19, True), (24, False)])
def test_persistent_kernel_equals_launch_chain_and_oracle(ntraj, cons):
    """the persistent kernels (one launch per run: the cooperative one with a grid barrier per step for one or two trajectories, the
    ensemble one -- eight trajectories per CTA, no grid barrier -- above that; config-1/2 shape: 603 dofs, two ml = 1 baths, fixed
    ends) against the per-step launch chain and the oracle: several run() calls, observables, a restart in between"""
    from sclmd_b200.engine import MDEngine
    c = P.md_case_c1_shape()
    K = P.psd_project(c["K"])
    nph, dt, nmd = K.shape[0], c["dt"], c["nmd"]
    cn = c["cons"] if cons else None
    kern = [np.array([np.diag(c["e"]["efric"][b])]) for b in range(2)]          # diagonal, ml = 1
    nz = [P.injected_noise(ntraj, nmd, 150, seed=90 + b, sigma=0.003) for b in range(2)]
    rng = np.random.default_rng(5)
    q0, p0 = 0.05 * rng.standard_normal((ntraj, nph)), 0.02 * rng.standard_normal((ntraj, nph))
    if cn:
        for g in cn:
            q0[:, g] = 0
            p0[:, g] = 0
    ens = O.EnsembleMD(K, dt, nmd, ntraj, cn)
    engs = []
    for persist in (True, False):
        e = MDEngine(nph, ntraj, dt, nmd)
        e.set_persistent(persist)
        e.set_dyn(K)
        if cn:
            e.set_constraint([i for g in cn for i in g])
        for b in range(2):
            e.add_bath(c["cids"][b], kern[b])
            e.set_noise(b, nz[b])
        e.set_state(q0, p0, 0)
        engs.append(e)
    for b in range(2):
        ens.add_bath(c["cids"][b], kern[b], nz[b])
    ens.q[:], ens.p[:] = q0, p0
    done = 0
    for chunk in (1, 37, 90):            # 128 steps in total: the slot indices wrap (nmd = 64)
        ens.run(chunk)
        done += chunk
        for e in engs:
            e.run(chunk)
            q, p, t = e.get_state()
            assert t == done and relerr(q, ens.q) < TOL_STEP and relerr(p, ens.p) < TOL_STEP, (done,)
    assert engs[0].launch_count() < 40 < engs[1].launch_count()                 # one launch per run() vs several per step
    for e in engs:
        assert relerr(e.etot(), ens.etot) < TOL_STEP
        for b in range(2):
            assert relerr(e.current(b), ens.baths[b]["cur"]) < TOL_OBS
    # hand the state of the chain engine to the persistent one and continue: both agree
    q, p, t = engs[1].get_state()
    engs[0].set_state(q, p, t)
    for e in engs:
        e.run(11)
    qa, pa, _ = engs[0].get_state()
    qb, pb, _ = engs[1].get_state()
    assert relerr(qa, qb) < 1e-12 and relerr(pa, pb) < 1e-12
    for e in engs:
        e.close()


@pytest.mark.parametrize("cons", [True, False])
def test_ensemble_kernel_with_a_dense_bath_equals_launch_chain_and_oracle(cons):
    """config-4 shape (726 dofs, two scalar-friction electron baths and the biased 36-dof bath with dense friction and
    non-conservative matrices, all time-local) as an ensemble of 11 trajectories: the ensemble-persistent kernel (dense bath:
    per-trajectory matrix-vector products inside the kernel, evaluations B and C not chained per element) against the per-step
    launch chain and the oracle"""
    from sclmd_b200.engine import MDEngine
    c = P.md_case_c4_shape()
    K = P.psd_project(c["K"])
    nph, dt, nmd, ntraj = K.shape[0], c["dt"], c["nmd"], 11
    cn = c["cons"] if cons else None
    lam = P.c4_lambda()
    damp = 100 / 0.658211814201041
    kern = [np.full((1, 120), 1.0 / damp), np.full((1, 120), 1.0 / damp), np.array([lam["eta_r"]])]
    bias = 0.7                                                               # baths.py:245-249: the q- and p-dependent extra forces
    Mq = [None, None, bias * (lam["xim_r"] - lam["zeta1_r"])]
    Mp = [None, None, -bias * lam["zeta2_r"]]
    nz = [P.injected_noise(ntraj, nmd, len(c["cids"][b]), seed=70 + b, sigma=0.003) for b in range(3)]
    rng = np.random.default_rng(6)
    q0, p0 = 0.05 * rng.standard_normal((ntraj, nph)), 0.02 * rng.standard_normal((ntraj, nph))
    if cn:
        for g in cn:
            q0[:, g] = 0
            p0[:, g] = 0
    ens = O.EnsembleMD(K, dt, nmd, ntraj, cn)
    engs = []
    for persist in (True, False):
        e = MDEngine(nph, ntraj, dt, nmd)
        e.set_persistent(persist)
        e.set_dyn(K)
        if cn:
            e.set_constraint([i for g in cn for i in g])
        for b in range(3):
            e.add_bath(c["cids"][b], kern[b], Mq[b], Mp[b])
            e.set_noise(b, nz[b])
        e.set_state(q0, p0, 0)
        engs.append(e)
    for b in range(3):
        if b < 2:
            ens.add_bath(c["cids"][b], kern[b], nz[b])
        else:
            ens.add_bath(c["cids"][b], kern[b], nz[b], bias=bias, exim=lam["xim_r"], zeta1=lam["zeta1_r"], zeta2=lam["zeta2_r"], kind="e")
    ens.q[:], ens.p[:] = q0, p0
    done = 0
    for chunk in (9, 40):                # 49 steps: the slot indices wrap (nmd = 32)
        ens.run(chunk)
        done += chunk
        for e in engs:
            e.run(chunk)
            q, p, t = e.get_state()
            assert t == done and relerr(q, ens.q) < TOL_STEP and relerr(p, ens.p) < TOL_STEP, (done,)
    assert engs[0].launch_count() < 40 < engs[1].launch_count()
    for e in engs:
        assert relerr(e.etot(), ens.etot) < TOL_STEP
        for b in range(3):
            assert relerr(e.current(b), ens.baths[b]["cur"]) < TOL_OBS
        e.close()
