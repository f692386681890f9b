"""Coloured-noise generation on the device vs the reference's golden series and the oracle."""
import os

import numpy as np
import pytest

import problems as P
from oracle import sclmd_oracle as O

pytestmark = pytest.mark.gpu
DT = 0.25 / 0.658


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def indexed_draws(av, stream):
    """the reference consumes one normal per strictly positive eigenvalue, in order (noise.py:299-303)"""
    xi = np.zeros(av.shape)
    k = 0
    for i in range(av.shape[0]):
        for j in range(av.shape[1]):
            if av[i, j] > 0:
                xi[i, j] = stream[k]
                k += 1
    return xi, k


def test_golden_series_with_reference_eigenvectors(golden_dir):
    """deterministic parity: the reference's own eigen-factors + its own normal draws -> its own series"""
    from sclmd_b200 import noise as N
    g = np.load(os.path.join(golden_dir, "noise.npz"))
    z = np.random.default_rng(60).standard_normal(4096)
    nmd = 32
    gwl, gam = P.gamma_grid(7, 4, 61, wmax=0.3)
    plan = N.ph_plan(gam, gwl, 300.0, float(g["phcut"]), DT, nmd)
    L = g["ph_au"] * np.sqrt(np.where(g["ph_av"] > 0, g["ph_av"], 0.0))[:, None, :]
    plan.set_factors(L)
    xi, used = indexed_draws(g["ph_av"], z)
    assert used == int(g["used_ph"])
    out = plan.generate(1, xi=xi[None])[0]
    assert relerr(out, np.real(g["ph"])) < 1e-12
    plan.close()
    efric, exim, exip = P.psd(3, 62, 0.05), P.antisym(3, 63, 0.01), P.sym(3, 64, 0.01)
    plan = N.e_plan(efric, exim, exip, 0.2, 300.0, 2.0, DT, nmd, False, False)
    assert plan.is_complex
    L = g["e_au"] * np.sqrt(np.where(g["e_av"] > 0, g["e_av"], 0.0))[:, None, :]
    plan.set_factors(L)
    xi, used = indexed_draws(g["e_av"], z[1000:])
    assert used == int(g["used_e"])
    out = plan.generate(1, xi=xi[None])[0]
    assert relerr(out, np.real(g["en"])) < 1e-12
    plan.close()


@pytest.mark.parametrize("nc,ngw,nmd", [(4, 7, 32), (9, 5, 64), (150, 3, 8), (170, 3, 6)])
def test_device_factor_reproduces_clamped_covariance(nc, ngw, nmd):
    """L L^H == clamp+(A(w)) for every frequency (noise.py:82-84 + vargau clamp)"""
    from sclmd_b200 import noise as N
    gwl, gam = P.gamma_grid(ngw, nc, 5, wmax=0.3)
    gam = gam - 0.004 * np.eye(nc)[None] * (np.arange(ngw) % 2)[:, None, None]      # make some nodes indefinite
    phcut = 0.8 * np.pi / DT
    plan = N.ph_plan(gam, gwl, 300.0, phcut, DT, nmd)
    L = plan.factors()
    for i in range(nmd // 2 + 1):
        A = O.ph_covariance(i, gam, gwl, 300.0, phcut, DT, nmd)
        ev, evec = np.linalg.eigh(A)
        want = (evec * np.where(ev > 0, ev, 0.0)) @ evec.T
        scale = max(np.abs(A).max(), 1e-300)
        assert np.abs(L[i] @ L[i].T - want).max() / scale < 1e-11, i
    plan.close()


def test_pivoted_cholesky_takes_semidefinite_spectra_and_jacobi_the_indefinite_ones():
    """positive SEMI-definite covariances (full rank, rank-deficient, zero above the cut-off) are factorised by the pivoted Cholesky
    kernel -- any L with L L^H = A gives the reference's Gaussian law (noise.py:82-84) -- and only the frequencies whose covariance
    has a negative eigenvalue (where vargau's clamp matters, noise.py:299-303) fall back to one-sided Jacobi"""
    from sclmd_b200 import noise as N
    nc, nmd = 40, 32
    nw = nmd // 2 + 1
    rng = np.random.default_rng(17)
    low = rng.standard_normal((nc, 5))
    gam = np.array([0.02 * low @ low.T / 5, P.psd(nc, 3, 0.02), P.psd(nc, 4, 0.02) - 0.004 * np.eye(nc)])   # rank 5 | full rank | indefinite
    gwl = np.array([0.0, 0.9, 1.8]) * np.pi / DT / 2
    phcut = 0.8 * np.pi / DT
    plan = N.ph_plan(gam, gwl, 300.0, phcut, DT, nmd)
    prof = plan.profile()
    assert prof["n_cholesky"] + prof["n_jacobi"] == nw and 0 < prof["n_jacobi"] < nw and prof["n_cholesky"] >= nw // 2
    L = plan.factors()
    for i in range(nw):
        A = O.ph_covariance(i, gam, gwl, 300.0, phcut, DT, nmd)
        ev, evec = np.linalg.eigh(A)
        want = (evec * np.where(ev > 0, ev, 0.0)) @ evec.T
        assert np.abs(L[i] @ L[i].T - want).max() / max(np.abs(A).max(), 1e-300) < 1e-11, i
    plan.close()
    # complex Hermitian PSD: efric dominates the bias terms
    efric, exim, exip = P.psd(12, 1, 0.05) + 0.05 * np.eye(12), P.antisym(12, 2, 0.002), P.sym(12, 3, 0.002)
    plan = N.e_plan(efric, exim, exip, 0.05, 300.0, 2.0, DT, nmd, False, True)
    assert plan.is_complex and plan.profile()["n_jacobi"] == 0
    L = plan.factors()
    for i in range(nw):
        A = O.e_covariance(i, efric, exim, exip, 0.05, 300.0, 2.0, DT, nmd, False, True)
        assert np.abs(L[i] @ L[i].conj().T - A).max() / max(np.abs(A).max(), 1e-300) < 1e-12, i
    plan.close()


def test_complex_factor_and_single_basis_shortcut():
    from sclmd_b200 import noise as N
    nc, nmd = 12, 16
    efric, exim, exip = P.psd(nc, 1, 0.05), P.antisym(nc, 2, 0.02), P.sym(nc, 3, 0.02)
    plan = N.e_plan(efric, exim, exip, 0.3, 300.0, 2.0, DT, nmd, False, False)
    L = plan.factors()
    for i in range(nmd // 2 + 1):
        A = O.e_covariance(i, efric, exim, exip, 0.3, 300.0, 2.0, DT, nmd, False, False)
        ev, evec = np.linalg.eigh(A)
        want = (evec * np.where(ev > 0, ev, 0.0)) @ evec.conj().T
        assert np.abs(L[i] @ L[i].conj().T - want).max() / max(np.abs(A).max(), 1e-300) < 1e-11, i
    plan.close()
    # efric only -> one eigendecomposition scaled per frequency, including an indefinite efric
    ef = P.sym(nc, 4, 0.05)
    plan = N.e_plan(ef, np.zeros((nc, nc)), np.zeros((nc, nc)), 0.0, 300.0, 2.0, DT, nmd)
    L = plan.factors()
    for i in range(nmd // 2 + 1):
        A = O.e_covariance(i, ef, np.zeros((nc, nc)), np.zeros((nc, nc)), 0.0, 300.0, 2.0, DT, nmd)
        ev, evec = np.linalg.eigh(A)
        want = (evec * np.where(ev > 0, ev, 0.0)) @ evec.T
        assert np.abs(L[i] @ L[i].T - want).max() / max(np.abs(A).max(), 1e-300) < 1e-11, i
    plan.close()


@pytest.mark.parametrize("nmd,nc,cplx", [(64, 5, False), (4096, 6, False), (8192, 4, False), (2000, 3, True),
                                         (20000, 2, False), (8192, 3, True)])
def test_transform_pipeline_vs_oracle(nmd, nc, cplx):
    """x = L xi, mirror, FFT/(dt nmd), real part -- against numpy with the SAME factors and draws;
    covers the in-shared-memory transform, the four-step path, radix 5 and odd column counts"""
    from sclmd_b200 import noise as N
    ntraj = 3
    if cplx:
        plan = N.e_plan(P.psd(nc, 1, 0.05), P.antisym(nc, 2, 0.02), P.sym(nc, 3, 0.02), 0.3, 300.0, 2.0, DT, nmd, False, False)
    else:
        gwl, gam = P.gamma_grid(4, nc, 8, wmax=0.3)
        plan = N.ph_plan(gam, gwl, 300.0, 2.0, DT, nmd)
    L = plan.factors()
    xi = np.random.default_rng(3).standard_normal((ntraj, nmd // 2 + 1, nc))
    out = plan.generate(ntraj, xi=xi)
    for k in range(ntraj):
        want = O.noise_from_factors(L, xi[k], DT, nmd)
        assert relerr(out[k], want) < 1e-11, k
    plan.close()


@pytest.mark.parametrize("nmd,nc", [(256, 5), (8192, 6), (20000, 3)])
def test_diagonal_spectra_skip_the_product_and_equal_the_general_path(nmd, nc, monkeypatch):
    """a diagonal friction spectrum (the diagonal memory kernels of config 5) has diagonal factors: the generator then multiplies the
    draws by the diagonals (k_fill_x_diag) instead of running x = L xi -- same Philox counters, the same series bit for bit as the
    general path (SCLMD_NOISE_NO_DIAG=1), with injected draws too, and equal to the oracle"""
    from sclmd_b200 import noise as N
    ntraj = 5
    gam = np.array([np.diag(np.linspace(0.01, 0.05, nc))])
    plan = N.ph_plan(gam, np.array([0.0]), 300.0, 2.0, DT, nmd)
    plan.generate(1, seed=1)                                    # (the first call also builds the twiddle / permutation tables)
    n0 = plan.launch_count()
    a = plan.generate(ntraj, seed=5, traj0=3)
    n_fast = plan.launch_count() - n0
    xi = np.random.default_rng(4).standard_normal((ntraj, nmd // 2 + 1, nc))
    ai = plan.generate(ntraj, xi=xi)
    monkeypatch.setenv("SCLMD_NOISE_NO_DIAG", "1")
    n0 = plan.launch_count()
    b = plan.generate(ntraj, seed=5, traj0=3)
    n_general = plan.launch_count() - n0
    bi = plan.generate(ntraj, xi=xi)
    monkeypatch.delenv("SCLMD_NOISE_NO_DIAG")
    assert np.array_equal(a, b) and np.array_equal(ai, bi)
    assert n_fast < n_general                                   # no product launch on the diagonal path
    L = plan.factors()
    assert not np.any(L * (1 - np.eye(nc))[None])               # the factors are diagonal
    for k in range(ntraj):
        assert relerr(ai[k], O.noise_from_factors(L, xi[k], DT, nmd)) < 1e-11, k
    # injected (general) factors switch the shortcut off
    rng = np.random.default_rng(9)
    Lg = L + 1e-3 * rng.standard_normal(L.shape) * np.sqrt(np.abs(L).max())
    plan.set_factors(Lg)
    ci = plan.generate(ntraj, xi=xi)
    for k in range(ntraj):
        assert relerr(ci[k], O.noise_from_factors(Lg, xi[k], DT, nmd)) < 1e-11, k
    plan.close()


def test_philox_draws_are_standard_normal_and_reproducible():
    from sclmd_b200 import noise as N
    nc, nmd, ntraj = 8, 4096, 16
    gam = np.array([np.eye(nc) * 0.01])
    plan = N.ph_plan(gam, np.array([0.0]), 300.0, 10.0, DT, nmd, classical=True)
    a = plan.generate(ntraj, seed=11)
    b = plan.generate(ntraj, seed=11)
    c = plan.generate(ntraj, seed=12)
    d = plan.generate(4, seed=11, traj0=4)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert np.array_equal(a[4:8], d)                       # counter = global trajectory index
    # classical white spectrum: A = dt nmd 2 kB T gamma ; series variance per sample = sum_w |x_w|^2 (2-sided)/(dt nmd)^2
    var_expected = 2 * O.KB * 300.0 * 0.01 / DT * (2 * (nmd // 2 - 1) + 2) / nmd
    v = a.var()
    assert abs(v / var_expected - 1) < 0.03
    # real spectrum -> even series, exactly as the reference produces (noise.py:82-94)
    assert relerr(a[:, 1:, :], a[:, :0:-1, :]) < 1e-10
    # different trajectories and dofs are uncorrelated
    flat = a.reshape(ntraj, -1)
    cc = np.corrcoef(flat)
    assert np.abs(cc - np.eye(ntraj)).max() < 0.05
    plan.close()


def test_gamt_vs_reference_golden(golden_dir):
    from sclmd_b200.baths import gamt
    g = np.load(os.path.join(golden_dir, "scalars.npz"))
    gwl, gam = P.gamma_grid(6, 3, 70, wmax=0.25)
    wl = [0.3 * i / 40 for i in range(40)]
    tl = [0.38 * i for i in range(9)]
    assert relerr(gamt(tl, wl, gwl, gam, 0), g["gamt0"]) < 1e-12
    assert relerr(gamt(tl, wl, gwl, gam, 0.01), g["gamt1"]) < 1e-12      # artificial damping branch (baths.py:43-50)
    # diagonal gamma, odd sizes
    out = gamt(tl[:5], wl[:33], gwl, gam[:, :1, :1], 0)
    assert relerr(out, O.gamt(tl[:5], wl[:33], gwl, gam[:, :1, :1], 0)) < 1e-12


def test_gmem_with_artificial_damping_rederives_gamma():
    """phbath.gmem with eta_ad != 0 (baths.py:429-445) against the reference formula evaluated with numpy"""
    from sclmd_b200.baths import phbath
    gwl, gam = P.gamma_grid(6, 3, 70, wmax=0.25)
    b = phbath(300.0, [0, 1, 2], 0.06, 40, DT, 64, ml=9, gamma=gam, gwl=gwl, eta_ad=0.01)
    b.gmem()
    tl = [DT * i for i in range(9)]
    kern = O.gamt(tl, b.wl, gwl, gam, 0.01)
    assert relerr(b.kernel, kern) < 1e-12
    want = np.zeros(gam.shape)
    for i in range(len(gwl)):
        for it in range(9):
            want[i] += DT * kern[it] * np.cos(gwl[i] * tl[it])
    assert relerr(b.gamma, want) < 1e-12 and relerr(b.gammaOld, gam) == 0


def test_gmem_from_a_self_energy_equals_the_reference(golden_dir):
    """phbath(sig=..., gwl=...) -> ggamma -> gmem (baths.py:322-327,375-395,412-446), both eta_ad branches, against the
    reference's own kernel and re-derived gamma (tests/golden/phbath_sig.npz)"""
    from sclmd_b200.baths import phbath
    g = np.load(os.path.join(golden_dir, "phbath_sig.npz"))
    nc, gwl, sig = P.phbath_sig_inputs()
    for tag, eta in (("eta0", 0), ("eta1", 0.02)):
        b = phbath(300.0, list(range(nc)), 0.06, 48, DT, 64, ml=11, mcof=2.0, sig=sig, gwl=gwl, eta_ad=eta)
        assert np.array_equal(b.gamma, g["gamma_" + tag])
        b.gmem()
        assert relerr(b.kernel, g["kernel_" + tag]) < 1e-12, tag
        assert relerr(b.gamma, g["gamma_after_" + tag]) < 1e-12, tag


def test_odd_nmd_is_rejected():
    from sclmd_b200 import noise as N
    from sclmd_b200._lib import SclmdError
    with pytest.raises(SclmdError):
        N.ph_plan(np.array([np.eye(2)]), np.array([0.0]), 300.0, 1.0, DT, 33)
    plan = N.ph_plan(np.array([np.eye(2)]), np.array([0.0]), 300.0, 1.0, DT, 2 * 7 * 11 * 13)   # not of the form 2^a 3^b 5^c
    with pytest.raises(SclmdError):
        plan.generate(1, seed=1)


def test_config4_junction_bath_with_example_matrices(golden_dir):
    """biased electron bath of examples/current-induced/rundp.py:76-77 (36x36 eta_r / xim_r / xip_r, bias 1.0, wmax 2.0,
    zpmotion False): the reference's series from its own eigen-factors and draws, and the device's own factors"""
    from sclmd_b200 import noise as N
    g = np.load(os.path.join(golden_dir, "noise_c4.npz"))
    lam = P.c4_lambda()
    dt, nmd = 0.5 / 0.658, 16
    z = np.random.default_rng(66).standard_normal(4096)
    plan = N.e_plan(lam["eta_r"], lam["xim_r"], lam["xip_r"], 1.0, 300.0, 2.0, dt, nmd, False, False)
    assert plan.is_complex
    Ldev = plan.factors()
    for i in range(nmd // 2 + 1):
        A = O.e_covariance(i, lam["eta_r"], lam["xim_r"], lam["xip_r"], 1.0, 300.0, 2.0, dt, nmd, False, False)
        ev, evec = np.linalg.eigh(A)
        want = (evec * np.where(ev > 0, ev, 0.0)) @ evec.conj().T
        assert np.abs(Ldev[i] @ Ldev[i].conj().T - want).max() / max(np.abs(A).max(), 1e-300) < 1e-10, i
    L = g["e_au"] * np.sqrt(np.where(g["e_av"] > 0, g["e_av"], 0.0))[:, None, :]
    plan.set_factors(L)
    xi, used = indexed_draws(g["e_av"], z)
    assert used == int(g["used"])
    out = plan.generate(1, xi=xi[None])[0]
    assert relerr(out, np.real(g["en"])) < 1e-12
    plan.close()


@pytest.mark.parametrize("n", [8, 60, 1000, 4096, 6000, 8192, 20000])
def test_myfft_and_power_spectrum_on_the_device(n):
    """functions.myfft (functions.py:11-53) and powerspecp (functions.py:221-236) with the in-house radix-2/3/4/5 transform,
    direct (n <= 6400) and four-step, against numpy.fft with the reference's normalisation"""
    from sclmd_b200.functions import myfft, powerspecp, device_fft
    rng = np.random.default_rng(n)
    dt = 0.25 / 0.658
    a = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    f = myfft(dt, n)
    dw = 2 * np.pi / dt / n
    assert relerr(f.iFourier1D(a), np.fft.fft(a) * dw / 2 / np.pi) < 1e-12
    assert relerr(f.Fourier1D(a), np.fft.ifft(a) * 2 * np.pi / dw) < 1e-12
    assert relerr(f.Fourier1D(f.iFourier1D(a)), a) < 1e-12                    # round trip
    b = rng.standard_normal((3, n))                                            # real batched input
    assert relerr(device_fft(b, -1, 1.0), np.fft.fft(b, axis=1)) < 1e-12
    if n <= 4096:
        ps = rng.standard_normal((n, 5))
        want = np.fft.ifft(ps.T, axis=1) * (2 * np.pi / dw)
        want = np.sum(np.real(want * np.conj(want)), axis=0) / dt / n
        got = powerspecp(ps, dt, n)
        assert np.array_equal(got[:, 0], np.arange(n) * dw) and relerr(got[:, 1], want) < 1e-11


def test_fft_rejects_unsupported_lengths():
    from sclmd_b200.functions import device_fft
    from sclmd_b200._lib import SclmdError
    with pytest.raises(SclmdError):
        device_fft(np.ones(14), -1, 1.0)          # 7 is not a supported radix


def test_vargau_reproduces_the_reference_draw_order():
    """noise.vargau (noise.py:273-305): deviates from np.random.normal drawn only for positive eigenvalues, V.r on the device"""
    from sclmd_b200.noise import vargau
    rng = np.random.default_rng(12)
    a = rng.standard_normal((7, 7))
    lam, vec = np.linalg.eigh(a @ a.T - 2.0 * np.eye(7))          # some negative eigenvalues
    assert (lam <= 0).any() and (lam > 0).any()
    np.random.seed(5)
    got = vargau(lam, vec, cof=0.5)
    np.random.seed(5)
    r = np.array([np.random.normal(0.0, np.sqrt(0.5 * v)) if v > 0 else 0.0 for v in lam])
    assert np.max(np.abs(got - vec @ r)) < 1e-12 * max(1.0, np.abs(vec @ r).max())
    h = a + 1j * rng.standard_normal((7, 7))
    lamc, vecc = np.linalg.eigh(h @ h.conj().T)
    np.random.seed(6)
    gotc = vargau(lamc, vecc)
    np.random.seed(6)
    rc = np.array([np.random.normal(0.0, np.sqrt(v)) if v > 0 else 0.0 for v in lamc])
    assert np.max(np.abs(gotc - vecc @ rc)) < 1e-12 * np.abs(vecc @ rc).max()
