"""The BASELINE.json configurations at their REAL sizes (SURVEY.md section 8, table of configs): the current-induced
example's noise length nmd = 2*10^5 (examples/current-induced/rundp.py:43), the full-kernel variant of config 5
(nc = 300, ml = 4096) and the per-frequency spectral factors at nc = 300 (noise.py:82,189).  CUDA path through the
C ABI against the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

import problems as P
from oracle import sclmd_oracle as O

pytestmark = pytest.mark.gpu


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


C4_DT, C4_NMD = 0.5 / 0.658, 200000          # rundp.py:41-43: dt = 0.5/0.658, nmd = 2*10^5 = 2^6 5^5


def test_config4_biased_bath_noise_at_full_length():
    """the biased 36-dof electron bath of rundp.py:76-77 at the example's own nmd: 100 001 complex Hermitian 36 x 36
    factors (spot-checked against eigh), then x = L xi, mirror, four-step radix-2/5 transform of length 200 000 for two
    trajectories with injected draws -- against numpy with the same factors and draws (noise.py:149-206)"""
    from sclmd_b200 import noise as N
    lam = P.c4_lambda()
    nw, nc, ntraj = C4_NMD // 2 + 1, 36, 2
    plan = N.e_plan(lam["eta_r"], lam["xim_r"], lam["xip_r"], 1.0, 300.0, 2.0, C4_DT, C4_NMD, False, False)
    assert plan.is_complex
    L = plan.factors()
    for i in (0, 1, 2, 777, 31250, 50000, 99999, 100000):
        A = O.e_covariance(i, lam["eta_r"], lam["xim_r"], lam["xip_r"], 1.0, 300.0, 2.0, C4_DT, C4_NMD, False, False)
        ev, evec = np.linalg.eigh(A)
        want = (evec * np.where(ev > 0, ev, 0.0)) @ evec.conj().T
        assert np.abs(L[i] @ L[i].conj().T - want).max() / max(np.abs(A).max(), 1e-300) < 1e-10, i
    xi = np.random.default_rng(21).standard_normal((ntraj, nw, nc))
    out = plan.generate(ntraj, xi=xi)
    assert out.shape == (ntraj, C4_NMD, nc)
    for k in range(ntraj):
        assert relerr(out[k], O.noise_from_factors(L, xi[k], C4_DT, C4_NMD)) < 1e-10, k
    prof = plan.profile()
    assert prof["transform_ms"] > 0 and prof["factor_ms"] > 0
    plan.close()


def test_config4_lead_bath_noise_at_full_length():
    """the two 120-dof electron baths of rundp.py:53-75 (efric = I/damp, no exim / exip: one basis matrix scaled per
    frequency) at nmd = 2*10^5, three trajectories with injected draws.  The factor of a multiple of the identity is
    L_w = sqrt(a_w/damp) I, so the expected series needs no 11 GB factor table on the host"""
    from sclmd_b200 import noise as N
    nw, nc, ntraj = C4_NMD // 2 + 1, 120, 3
    damp = 100 / 0.658211814201041
    ef, z = np.identity(nc) / damp, np.zeros((nc, nc))
    plan = N.e_plan(ef, z, z, 0.0, 300.0, 2.0, C4_DT, C4_NMD, False, False)
    assert not plan.is_complex
    xi = np.random.default_rng(22).standard_normal((ntraj, nw, nc))
    out = plan.generate(ntraj, xi=xi)
    aw = np.array([O.e_coefficients(i, 0.0, 300.0, 2.0, C4_DT, C4_NMD, False, False)[0] for i in range(nw)])
    amp = np.sqrt(np.where(aw > 0, aw, 0.0) / damp)
    for k in range(ntraj):
        want = np.real(O.spectrum_to_series(amp[:, None] * xi[k], C4_DT, C4_NMD))
        assert relerr(out[k], want) < 1e-10, k
    # the device's own Philox draws: reproducible, and independent of how the ensemble is cut into trajectory blocks
    a = plan.generate(2, seed=5, traj0=3)
    b = plan.generate(1, seed=5, traj0=4)
    assert np.array_equal(a[1], b[0])
    plan.close()


def test_myfft_at_the_config4_length():
    """functions.myfft (functions.py:11-53) at N = 2*10^5: four-step transform with radix-2/4/5 passes"""
    from sclmd_b200.functions import myfft
    n = C4_NMD
    rng = np.random.default_rng(n)
    a = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    f = myfft(C4_DT, n)
    dw = 2 * np.pi / C4_DT / n
    assert relerr(f.iFourier1D(a), np.fft.fft(a) * dw / 2 / np.pi) < 1e-11
    assert relerr(f.Fourier1D(f.iFourier1D(a)), a) < 1e-11


@pytest.mark.parametrize("ntraj", [3, 130])
def test_config5_full_kernels_at_full_size_vs_oracle(ntraj):
    """BASELINE configs[4], full-kernel variant, per-trajectory shape at full size: 3000 dofs, two baths of 300 dofs with
    FULL 300 x 300 memory kernels of 4096 steps (2.9 GB per bath), random pre-existing history so that the whole memory acts
    from the first step.  ntraj = 3 takes the skinny cp.async GEMM tiles, ntraj = 130 the production path (TMA-fed stream-K contraction
    over the rotating ring: 77 805 K-slabs per tile cut across all SMs, a ragged last row tile); <= 1e-10 per step against the oracle (baths.py:448-458)"""
    from sclmd_b200.engine import MDEngine
    natoms, nc, ml, nmd = 1000, 300, 4096, 32
    nph, dt = 3 * natoms, 0.25 / 0.658
    K = P.spring_chain_dyn(natoms, seed=5)
    cids = [list(range(0, nc)), list(range(nph - nc, nph))]
    rng = np.random.default_rng(51)
    q0, p0 = 0.05 * rng.standard_normal((ntraj, nph)), 0.02 * rng.standard_normal((ntraj, nph))
    ens = O.EnsembleMD(K, dt, nmd, ntraj, None)
    eng = MDEngine(nph, ntraj, dt, nmd)
    eng.set_dyn(K)
    eng.set_state(q0, p0, 0)
    for b in range(2):
        kern = P.full_kernel(ml, nc, dt, seed=30 + b, tau=600.0)
        nz = P.injected_noise(ntraj, nmd, nc, seed=40 + b)
        hist = 0.02 * rng.standard_normal((ntraj, ml, nc))          # phis[i] = p_{-1-i}[cids]
        ens.add_bath(cids[b], kern, nz)
        ens.baths[b]["ring"][:, (-1 - np.arange(ml)) % ml, :] = hist
        eng.add_bath(cids[b], kern)
        eng.set_noise(b, nz)
        eng.set_history(b, hist)
        del kern, hist
    ens.q[:], ens.p[:] = q0, p0
    ens.t = -1
    for b in ens.baths:
        b["tail"] = ens._tail(b)             # S(0) = dt sum_j k[j] p_{-j}
    ens.t = 0
    done = 0
    for chunk in (3, 4):
        ens.run(chunk)
        eng.run(chunk)
        done += chunk
        q, p, t = eng.get_state()
        assert t == done and relerr(q, ens.q) < 1e-10 and relerr(p, ens.p) < 1e-10, done
    for b in range(2):
        assert relerr(eng.current(b)[:, :done], ens.baths[b]["cur"][:, :done]) < 1e-8
    assert relerr(eng.etot()[:, :done], ens.etot[:, :done]) < 1e-10
    eng.close()


def test_spectral_factors_at_nc_300():
    """per-frequency factorisation at the config-5 bath size (noise.py:82: eigh of a 300 x 300 covariance per frequency) on a
    3-node gamma grid, i.e. a different matrix for every frequency: L L^T == clamp+(A(w)) (vargau's clamp, noise.py:299-303).
    300 x 300 does not fit shared memory: this is the global-memory path of the one-sided Jacobi kernel"""
    from sclmd_b200 import noise as N
    nc, nmd, dt = 300, 12, 0.25 / 0.658
    gwl, gam = P.gamma_grid(3, nc, 9, wmax=0.9 * np.pi / dt)
    gam[1] -= 0.002 * np.eye(nc)                                   # an indefinite node: the clamp matters
    phcut = 0.95 * np.pi / dt
    plan = N.ph_plan(gam, gwl, 300.0, phcut, dt, nmd)
    L = plan.factors()
    for i in range(nmd // 2 + 1):
        A = O.ph_covariance(i, gam, gwl, 300.0, phcut, dt, nmd)
        ev, evec = np.linalg.eigh(A)
        want = (evec * np.where(ev > 0, ev, 0.0)) @ evec.T
        assert np.abs(L[i] @ L[i].T - want).max() / max(np.abs(A).max(), 1e-300) < 1e-10, i
    xi = np.random.default_rng(3).standard_normal((2, nmd // 2 + 1, nc))
    out = plan.generate(2, xi=xi)
    for k in range(2):
        assert relerr(out[k], O.noise_from_factors(L, xi[k], dt, nmd)) < 1e-11
    plan.close()


def test_getters_order_themselves_after_an_asynchronous_run():
    """sclmd_md_run(h, n, NULL) only enqueues; every sclmd_md_get_* must synchronise with the handle's stream by itself
    (include/sclmd_b200.h) -- read the observables and the history FIRST, without a get_state in between"""
    from sclmd_b200.engine import MDEngine
    natoms, nc, ml, ntraj, nmd = 400, 60, 700, 64, 256
    nph, dt = 3 * natoms, 0.3
    K = P.psd_project(P.spring_chain_dyn(natoms, seed=3))
    kern = P.diag_kernel(ml, nc, dt, 1)
    nz = P.injected_noise(ntraj, nmd, nc, seed=2)

    def mk():
        e = MDEngine(nph, ntraj, dt, nmd)
        e.set_dyn(K)
        e.add_bath(list(range(3, 3 + nc)), kern)
        e.set_noise(0, nz)
        e.set_state(np.full((ntraj, nph), 0.01), np.zeros((ntraj, nph)), 0)
        return e
    a, b = mk(), mk()
    a.run(200)
    want = (a.current(0), a.etot(), a.get_history(0))
    b.run_async(200)
    cur = b.current(0)                   # no get_state before: the getter itself has to wait for the 200 steps
    b.set_state(np.full((ntraj, nph), 0.01), np.zeros((ntraj, nph)), 0)
    b.reset_history()
    b.run_async(200)
    et = b.etot()
    b.set_state(np.full((ntraj, nph), 0.01), np.zeros((ntraj, nph)), 0)
    b.reset_history()
    b.run_async(200)
    hist = b.get_history(0)
    assert np.array_equal(cur, want[0]) and np.array_equal(et, want[1]) and np.array_equal(hist, want[2])
    a.close()
    b.close()
