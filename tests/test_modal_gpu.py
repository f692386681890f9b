"""Eigenbasis ("modal") propagation of the CUDA integrator (sclmd_md_set_modes; specification oracle.ModalMD) against the
real-space oracle (oracle.EnsembleMD, itself pinned to the reference's md.vv): same trajectories, observables and histories to
rounding.  A 1-ulp change of K moves a real-space trajectory by the same amount (about 1e-15 per step, 1e-12 after 4096 steps),
so the per-step tolerance of the other tests applies unchanged."""
import numpy as np
import pytest

import problems as P
from oracle import sclmd_oracle as O

pytestmark = pytest.mark.gpu
TOL_STEP, TOL_OBS = 1e-10, 1e-8


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def modes(Kraw):
    """what md.setDyn keeps (md.py:264-281): clipped eigenvalues, eigenvectors, the rebuilt matrix"""
    lam, U = np.linalg.eigh(0.5 * (Kraw + Kraw.T))
    lam = np.where(lam < 0, 0.0, lam)
    return U @ np.diag(lam) @ U.T, lam, U


def build(natoms, baths, ntraj, nmd, dt, seed, modal=True, cons=None):
    """engine (+ modes) and oracle for `baths` = [(cids, ml)] with diagonal kernels"""
    from sclmd_b200.engine import MDEngine
    nph = 3 * natoms
    K, lam, U = modes(P.spring_chain_dyn(natoms, seed=seed))
    eng = MDEngine(nph, ntraj, dt, nmd)
    eng.set_dyn(K)
    eng.set_modes(lam, U)
    eng.set_modal(modal)
    eng.set_persistent(False)          # small time-local cases would otherwise take the persistent kernels
    ens = O.EnsembleMD(K, dt, nmd, ntraj, cons)
    if cons:
        eng.set_constraint([i for g in cons for i in g])
    for b, (cids, ml) in enumerate(baths):
        kern = P.diag_kernel(ml, len(cids), dt, 50 + b, tau=40.0)
        nz = P.injected_noise(ntraj, nmd, len(cids), seed=60 + b)
        eng.add_bath(cids, kern)
        eng.set_noise(b, nz)
        ens.add_bath(cids, kern, nz)
    rng = np.random.default_rng(seed + 1)
    q0, p0 = 0.05 * rng.standard_normal((ntraj, nph)), 0.02 * rng.standard_normal((ntraj, nph))
    eng.set_state(q0, p0, 0)
    ens.q[:], ens.p[:] = q0, p0
    return eng, ens


def check_state(eng, ens, done):
    q, p, t = eng.get_state()
    assert t == done
    assert relerr(q, ens.q) < TOL_STEP and relerr(p, ens.p) < TOL_STEP, done


@pytest.mark.parametrize("natoms,baths,ntraj", [
    (40, [(list(range(0, 12)), 1), (list(range(108, 120)), 1)], 5),                 # time-local baths
    (40, [(list(range(0, 12)), 50), (list(range(100, 113)), 7)], 3),                # memory kernels (direct tails), odd nc
    (60, [([7, 3, 12, 4, 9, 30, 31], 300), (list(range(150, 180)), 1)], 70),       # unsorted dofs, time-blocked tail, > 64 trajectories
    (30, [(list(range(10, 40)), 140)], 1),                                         # one bath, one trajectory
])
def test_modal_propagation_vs_oracle(natoms, baths, ntraj):
    nmd, dt = 64, 0.25 / 0.658
    eng, ens = build(natoms, baths, ntraj, nmd, dt, seed=5)
    assert eng.modal_active()
    done = 0
    for chunk in (1, 2, 30, 17, 100):            # get_state in between: leaves and re-enters the eigenbasis; slots wrap (nmd = 64)
        eng.run(chunk)
        ens.run(chunk)
        done += chunk
        check_state(eng, ens, done)
    assert relerr(eng.etot(), ens.etot) < TOL_STEP
    for b in range(len(baths)):
        assert relerr(eng.current(b), ens.baths[b]["cur"]) < TOL_OBS
        ml = baths[b][1]
        want = np.stack([ens.baths[b]["ring"][:, (done - 1 - i) % ml, :] for i in range(ml)], axis=1)
        assert relerr(eng.get_history(b), want) < TOL_STEP
    eng.close()


def test_modal_equals_real_space_engine_and_survives_a_restart():
    """the same handle configuration with the eigenbasis switched off, asynchronous runs without reads in between, and
    state + history handed to a fresh modal engine in the middle"""
    natoms, ntraj, nmd, dt = 50, 9, 128, 0.3
    baths = [(list(range(3, 23)), 200), (list(range(120, 141)), 33)]
    a, ens = build(natoms, baths, ntraj, nmd, dt, seed=8, modal=True)
    b, _ = build(natoms, baths, ntraj, nmd, dt, seed=8, modal=False)
    assert a.modal_active() and not b.modal_active()
    for e in (a, b):
        e.run_async(40)
        e.run_async(1)
        e.run_async(59)
    ens.run(100)
    qa, pa, ta = a.get_state()
    qb, pb, tb = b.get_state()
    assert ta == tb == 100 and relerr(qa, qb) < 1e-12 and relerr(pa, pb) < 1e-12
    check_state(a, ens, 100)
    assert relerr(a.current(0), b.current(0)) < 1e-10 and relerr(a.etot(), b.etot()) < 1e-11
    c, _ = build(natoms, baths, ntraj, nmd, dt, seed=8, modal=True)
    c.set_state(qa, pa, ta)
    for i in range(2):
        c.set_history(i, a.get_history(i))
    for e in (a, c):
        e.run(37)
    ens.run(37)
    check_state(c, ens, 137)
    check_state(a, ens, 137)
    for e in (a, b, c):
        e.close()


def test_modal_streamed_noise_rows_match_resident_table():
    """the end-to-end path of bench.py in the eigenbasis: a noise row uploaded per step, one asynchronous step, observables read
    back without touching the state -- bit-identical to the same engine on a resident table"""
    from sclmd_b200.engine import MDEngine
    natoms, nc, ml, ntraj, dt, nmd = 30, 10, 150, 6, 0.3, 16
    nph = 3 * natoms
    K, lam, U = modes(P.spring_chain_dyn(natoms, seed=3))
    kern = P.diag_kernel(ml, nc, dt, 1)
    nz = P.injected_noise(ntraj, nmd, nc, seed=2)

    def mk():
        e = MDEngine(nph, ntraj, dt, nmd)
        e.set_dyn(K)
        e.set_modes(lam, U)
        e.add_bath(list(range(3, 3 + nc)), kern)
        e.set_state(np.full((ntraj, nph), 0.01), np.zeros((ntraj, nph)), 0)
        return e
    a, b = mk(), mk()
    assert a.modal_active()
    a.set_noise(0, nz)
    rows = np.ascontiguousarray(nz.transpose(1, 0, 2))
    b.set_noise_rows(0, 0, rows[0:1])
    nsteps = 40
    obs = []
    for t in range(nsteps):
        b.set_noise_rows(0, (t + 1) % nmd, rows[(t + 1) % nmd:(t + 1) % nmd + 1])
        b.run_async(1)
        obs.append(b.step_observables(t % nmd).copy())
    a.run_async(nsteps)
    qa, pa, ta = a.get_state()
    qb, pb, tb = b.get_state()
    assert ta == tb == nsteps and np.array_equal(qa, qb) and np.array_equal(pa, pb)
    et, cur = a.etot(), a.current(0)
    for t in range(nsteps - nmd, nsteps):
        assert np.array_equal(obs[t][0], et[:, t % nmd]) and np.array_equal(obs[t][1], cur[:, t % nmd])
    a.close()
    b.close()


def test_modal_falls_back_to_real_space_when_the_problem_does_not_allow_it():
    from sclmd_b200.engine import MDEngine
    from sclmd_b200._lib import SclmdError
    natoms, nmd, dt = 20, 16, 0.3
    nph = 3 * natoms
    K, lam, U = modes(P.spring_chain_dyn(natoms, seed=4))

    def mk():
        e = MDEngine(nph, 3, dt, nmd)
        e.set_dyn(K)
        e.set_modes(lam, U)
        return e
    e = mk()
    assert not e.modal_active()                              # no bath: nothing to gain
    e.add_bath(list(range(0, 6)), P.diag_kernel(4, 6, dt, 1))
    assert e.modal_active()
    e.set_constraint([57, 58, 59])
    assert not e.modal_active()                              # constraints act in real space
    e.set_constraint([])
    assert e.modal_active()
    e.add_bath(list(range(4, 9)), P.diag_kernel(2, 5, dt, 2))
    assert not e.modal_active()                              # overlapping baths
    e.close()
    e = mk()
    e.add_bath(list(range(0, 6)), P.full_kernel(3, 6, dt, 1))
    assert not e.modal_active()                              # full kernels couple the bath dofs
    e.close()
    e = mk()
    e.add_bath(list(range(0, 40)), P.diag_kernel(3, 40, dt, 1))
    assert not e.modal_active()                              # 2 sum(nc) > nph: K.q is cheaper
    e.set_dyn(K)
    assert not e.modal_active()
    with pytest.raises(SclmdError):                          # a decomposition that does not belong to K
        e.set_modes(lam * 1.001, U)
    e.set_modes(lam, U)
    e.close()
    # the fall-back really runs: constraints + modes against the oracle
    cons = [list(range(0, 3)), list(range(57, 60))]
    eng, ens = build(natoms, [(list(range(3, 13)), 9)], 4, nmd, dt, seed=4, cons=cons)
    assert not eng.modal_active()
    eng.run(25)
    ens.run(25)
    check_state(eng, ens, 25)
    eng.close()


def test_config5_shape_full_size_modal_vs_oracle():
    """BASELINE configs[4] per-trajectory shape at full size (3000 dofs, 2 x 300 bath dofs, diagonal 4096-step kernels), random
    pre-existing history: eigenbasis propagation with the time-blocked ring pass against the oracle across a block boundary"""
    from sclmd_b200.engine import MDEngine
    natoms, nc, ml, ntraj, nmd = 1000, 300, 4096, 4, 64
    nph, dt = 3 * natoms, 0.25 / 0.658
    K, lam, U = modes(P.spring_chain_dyn(natoms, seed=5))
    cids = [list(range(0, nc)), list(range(nph - nc, nph))]
    kern = [P.diag_kernel(ml, nc, dt, 30 + b, tau=600.0) for b in range(2)]
    nz = [P.injected_noise(ntraj, nmd, nc, seed=40 + b) for b in range(2)]
    rng = np.random.default_rng(50)
    q0, p0 = 0.05 * rng.standard_normal((ntraj, nph)), 0.02 * rng.standard_normal((ntraj, nph))
    hist = [0.02 * rng.standard_normal((ntraj, ml, nc)) for _ in range(2)]
    ens = O.EnsembleMD(K, dt, nmd, ntraj, None)
    for b in range(2):
        ens.add_bath(cids[b], kern[b], nz[b])
        ens.baths[b]["ring"][:, (-1 - np.arange(ml)) % ml, :] = hist[b]
    ens.q[:], ens.p[:] = q0, p0
    ens.t = -1
    for b in ens.baths:
        b["tail"] = ens._tail(b)
    ens.t = 0
    e = MDEngine(nph, ntraj, dt, nmd)
    e.set_dyn(K)
    e.set_modes(lam, U)
    for b in range(2):
        e.add_bath(cids[b], kern[b])
        e.set_noise(b, nz[b])
    e.set_state(q0, p0, 0)
    for b in range(2):
        e.set_history(b, hist[b])
    assert e.modal_active()
    done = 0
    for chunk in (20, 17):
        ens.run(chunk)
        e.run(chunk)
        done += chunk
        check_state(e, ens, done)
    for b in range(2):
        assert relerr(e.current(b)[:, :done], ens.baths[b]["cur"][:, :done]) < TOL_OBS
    assert relerr(e.etot()[:, :done], ens.etot[:, :done]) < TOL_STEP
    e.close()


def test_eigenbasis_step_on_one_stream_equals_the_two_stream_step():
    """sclmd_md_set_overlap(h, 0): the products, the modal update and the tail kernels of the eigenbasis step all on the handle's stream
    (A/B switch: measured 0.567 against 0.534 ms per step at config 5, the near passes are hidden beside the products otherwise)"""
    natoms, ntraj, nmd, dt = 40, 4, 128, 0.3
    baths = [(list(range(3, 23)), 160), (list(range(90, 111)), 1)]
    a, ens = build(natoms, baths, ntraj, nmd, dt, seed=21, modal=True)
    b, _ = build(natoms, baths, ntraj, nmd, dt, seed=21, modal=True)
    b.set_overlap(False)
    for e in (a, b):
        e.run_async(70)
    ens.run(70)
    qa, pa, _ = a.get_state()
    qb, pb, _ = b.get_state()
    assert np.array_equal(qa, qb) and np.array_equal(pa, pb)
    check_state(b, ens, 70)
    a.close()
    b.close()
