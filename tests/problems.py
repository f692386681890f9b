"""Seeded synthetic problem definitions shared by oracle/make_golden.py, the
parity tests and bench.py (SURVEY.md section 8d "Synthetic inputs").

Everything is a pure function of its seed so the golden fixtures only need to
store the reference's OUTPUTS; inputs are regenerated on both boxes.
"""
import numpy as np


from sclmd_b200.synthetic import (spring_chain_dyn, psd_project, psd_project_modes, full_kernel, diag_kernel)  # noqa: F401,E402


def injected_noise(ntraj, nmd, nc, seed, sigma=0.01):
    return sigma * np.random.default_rng(seed).standard_normal((ntraj, nmd, nc))


def sym(n, seed, scale=1.0):
    a = np.random.default_rng(seed).standard_normal((n, n))
    return scale * 0.5 * (a + a.T)


def antisym(n, seed, scale=1.0):
    a = np.random.default_rng(seed).standard_normal((n, n))
    return scale * 0.5 * (a - a.T)


def psd(n, seed, scale=1.0):
    a = np.random.default_rng(seed).standard_normal((n, n))
    return scale * (a @ a.T) / n


def gamma_grid(ngw, nc, seed, wmax=0.3):
    """A PSD friction spectrum gamma(w) on a uniform grid [0,wmax]: [ngw,nc,nc]."""
    gwl = np.linspace(0.0, wmax, ngw)
    A, B = psd(nc, seed, 0.02), psd(nc, seed + 1, 0.02)
    g = np.array([A * np.exp(-(w / 0.12) ** 2) + B * (w / wmax) * np.exp(-(w / 0.2) ** 2) for w in gwl])
    return gwl, g


def chain_blocks(m, seed=0, k=1.0, k2=0.15):
    """Principal-layer blocks of a quasi-1D harmonic chain (m dofs per layer):
    K00 = K11 on-site, K01 coupling layer0->layer1 (selfenergy.py:93-103).
    Units ps^-2-like, O(1e3) to mimic real dynamical matrices."""
    rng = np.random.default_rng(seed)
    # energy  sum_n |a u_n - b u_{n+1}|^2 / 2  + onsite  ->  K00 = a^T a + b^T b, K01 = -a^T b
    a = np.sqrt(1.0e3 * k) * (np.eye(m) + k2 * rng.uniform(-1.0, 1.0, (m, m)))
    b = np.sqrt(1.0e3 * k) * (np.eye(m) + k2 * rng.uniform(-1.0, 1.0, (m, m)))
    on = a.T @ a + b.T @ b + 5.0 * np.eye(m)
    on = 0.5 * (on + on.T)
    return on, on.copy(), -(a.T @ b)


# ---------------------------------------------------------------- MD parity cases
def md_case_ph_full():
    """two phonon baths with full memory kernels of different length + constraints"""
    natoms, dt, nmd, nsteps = 10, 0.25 / 0.658, 32, 40
    K = spring_chain_dyn(natoms, seed=11)
    cons = [list(range(0, 3)), list(range(27, 30))]
    cid = [list(range(3, 9)), list(range(21, 27))]
    mls = [5, 3]
    kern = [full_kernel(mls[b], 6, dt, seed=20 + b) for b in range(2)]
    noise = [injected_noise(1, nmd, 6, seed=30 + b)[0] for b in range(2)]
    rng = np.random.default_rng(40)
    q0, p0 = 0.05 * rng.standard_normal(30), 0.02 * rng.standard_normal(30)
    for c in cons:
        q0[c] = 0
        p0[c] = 0
    return dict(K=K, dt=dt, nmd=nmd, nsteps=nsteps, cons=cons, cids=cid, kern=kern, noise=noise,
                q0=q0, p0=p0, kinds=["ph", "ph"], T=300.0)


def md_case_ph_local():
    """Debye phonon baths (ml=1, no dt factor), no constraints"""
    natoms, dt, nmd, nsteps = 8, 0.25 / 0.658, 16, 20
    K = spring_chain_dyn(natoms, seed=12)
    cid = [list(range(0, 6)), list(range(18, 24))]
    debye = 0.05
    kern = [np.array([np.diag(debye * np.pi / 6.0 + np.zeros(6))]) for _ in range(2)]
    noise = [injected_noise(1, nmd, 6, seed=33 + b)[0] for b in range(2)]
    rng = np.random.default_rng(41)
    return dict(K=K, dt=dt, nmd=nmd, nsteps=nsteps, cons=None, cids=cid, kern=kern, noise=noise,
                q0=0.05 * rng.standard_normal(24), p0=0.02 * rng.standard_normal(24), kinds=["ph", "ph"], T=300.0)


def md_case_e_extra():
    """electron baths: one with exim/zeta1/zeta2 all set (extra forces act), one with
    zeta1=zeta2=None (baths.py:233 quirk: exim force silently skipped) + constraints"""
    natoms, dt, nmd, nsteps = 10, 0.5 / 0.658, 32, 36
    K = spring_chain_dyn(natoms, seed=13)
    cons = [list(range(0, 3)), list(range(27, 30))]
    cid = [list(range(3, 12)), list(range(15, 24))]
    noise = [injected_noise(1, nmd, 9, seed=36 + b)[0] for b in range(2)]
    e = dict(efric=[psd(9, 50, 0.03), psd(9, 51, 0.03)],
             exim=[antisym(9, 52, 0.01), antisym(9, 53, 0.01)],
             exip=[sym(9, 54, 0.01), sym(9, 55, 0.01)],
             zeta1=[sym(9, 56, 0.004), None], zeta2=[antisym(9, 57, 0.004), None],
             bias=[0.7, 0.4])
    rng = np.random.default_rng(42)
    q0, p0 = 0.05 * rng.standard_normal(30), 0.02 * rng.standard_normal(30)
    for c in cons:
        q0[c] = 0
        p0[c] = 0
    return dict(K=K, dt=dt, nmd=nmd, nsteps=nsteps, cons=cons, cids=cid, noise=noise, e=e,
                q0=q0, p0=p0, kinds=["e", "e"], T=300.0)


def md_case_c1_shape():
    """config-1 shape: 201 atoms, nph=603, fixed ends, two ml=1 baths on 150 dofs each with
    efric = I/damp (examples/runmd.py:31-54), random-phase normal-mode initial conditions."""
    natoms, dt, nmd, nsteps = 201, 0.25 / 0.658, 64, 24
    K = spring_chain_dyn(natoms, seed=14)
    cons = [list(range(0 * 3, 20 * 3)), list(range(181 * 3, 201 * 3))]
    cid = [list(range(20 * 3, 70 * 3)), list(range(131 * 3, 181 * 3))]
    damp = 100 / 0.658211814201041
    e = dict(efric=[np.identity(150) / damp] * 2, exim=[None] * 2, exip=[None] * 2, zeta1=[None] * 2,
             zeta2=[None] * 2, bias=[0.0, 0.0])
    noise = [injected_noise(1, nmd, 150, seed=38 + b, sigma=0.003)[0] for b in range(2)]
    return dict(K=K, dt=dt, nmd=nmd, nsteps=nsteps, cons=cons, cids=cid, noise=noise, e=e,
                q0=None, p0=None, kinds=["e", "e"], T=300.0, ic_seed=77)


def c4_lambda():
    """The 36x36 electron-phonon matrices of the reference's current-induced example (eta_r, xim_r, xip_r, zeta1_r,
    zeta2_r of examples/current-induced/grapheneLambda-r-0.3-ver2.nc), extracted into a fixture by oracle/make_golden.py
    with sclmd_b200.myio (the reference tree does not exist on the GPU box)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c4_lambda.npz"))
    return {k: np.array(g[k]) for k in g.files}


def md_case_c4_shape():
    """config-4 shape (examples/current-induced/rundp.py:51-79): 242 atoms, nph = 726, 24 + 48 fixed dofs, two electron
    baths on 120 dofs each with efric = I/damp, and a biased electron bath (bias = 1.0) on the 36 junction dofs with the
    example's own friction / non-conservative / Berry matrices; zeta1 = zeta2 = None as in the script, so the
    baths.py:233 quirk applies; dt = 0.5/0.658; zero initial conditions (noranvel); harmonic force constants."""
    natoms, dt, nmd, nsteps = 242, 0.5 / 0.658, 32, 12
    K = spring_chain_dyn(natoms, seed=15)
    cons = [list(range(0 * 3, (7 + 1) * 3)), list(range(226 * 3, (241 + 1) * 3))]
    cid = [list(range(8 * 3, (47 + 1) * 3)), list(range(186 * 3, (225 + 1) * 3)), list(range(111 * 3, (122 + 1) * 3))]
    damp = 100 / 0.658211814201041
    lam = c4_lambda()
    e = dict(efric=[np.identity(120) / damp, np.identity(120) / damp, lam["eta_r"]],
             exim=[None, None, lam["xim_r"]], exip=[None, None, lam["xip_r"]],
             zeta1=[None] * 3, zeta2=[None] * 3, bias=[0.0, 0.0, 1.0])
    noise = [injected_noise(1, nmd, len(cid[b]), seed=44 + b, sigma=0.003)[0] for b in range(3)]
    n = 3 * natoms
    return dict(K=K, dt=dt, nmd=nmd, nsteps=nsteps, cons=cons, cids=cid, noise=noise, e=e,
                q0=np.zeros(n), p0=np.zeros(n), kinds=["e", "e", "e"], T=300.0)


MD_CASES = dict(ph_full=md_case_ph_full, ph_local=md_case_ph_local, e_extra=md_case_e_extra, c1_shape=md_case_c1_shape,
                c4_shape=md_case_c4_shape)


def myio_inputs():
    """seeded contents of the three file kinds the reference's remaining NetCDF readers take (myio.py:192-366): a biased Lambda
    file (wl, muLR, ImPir2, RePir2, ReLamLR), a wide-band one (eta .. zeta2), a PHrun phonon file (hw, U, DynamicAtoms) and a
    self-energy file (Wlist, Re/ImSigL/R)"""
    rng = np.random.default_rng(77)
    n, nw, natoms_all, first, last = 6, 7, 5, 2, 3          # dynamic atoms 2..3 (1-based) of 5 -> 6 dofs
    lam = dict(wl=np.linspace(0.01, 0.19, nw), muLR=np.array([0.35, -0.25]), ImPir2=rng.standard_normal((nw, n, n)),
               RePir2=rng.standard_normal((nw, n, n)), ReLamLR=rng.standard_normal((nw, n, n)))
    wb = dict(eta=psd(n, 71, 0.1), xim=antisym(n, 72, 0.1), xip=sym(n, 73, 0.1), zeta1=sym(n, 74, 0.1), zeta2=antisym(n, 75, 0.1))
    ph = dict(hw=0.01 + 0.2 * rng.random(n), U=rng.standard_normal((n, natoms_all, 3)), DynamicAtoms=np.arange(first, last + 1, dtype=np.int32))
    sg = dict(Wlist=np.linspace(0, 0.2, nw), ReSigL=rng.standard_normal((nw, 3, 3)), ImSigL=rng.standard_normal((nw, 3, 3)),
              ReSigR=rng.standard_normal((nw, 3, 3)), ImSigR=rng.standard_normal((nw, 3, 3)))
    return dict(lam=lam, wb=wb, ph=ph, sg=sg)


# ---------------------------------------------------------------- post-processing / pipeline-glue cases
# name, seed, baths, runs, T, calHF / calTC arguments (tools.py:132-215)
KAPPA_CASES = [("b2_d1", 11, 2, 7, 300, dict(delta=0.1, dlist=1)),
               ("b2_d0", 12, 2, 5, 300, dict(delta=0.2, dlist=0)),
               ("b3_d2", 13, 3, 6, 250, dict(delta=0.1, dlist=2, L=12.5, A=30.0)),
               ("b2_delta0", 14, 2, 4, 300, dict(delta=0, dlist=1))]


def kappa_case(seed, bathnum, nruns, T):
    """seeded kappa.<T>.bath<i>.run<j>.dat contents as md.Run writes them (md.py:658-664): run index, T, mean current"""
    rng = np.random.default_rng(seed)
    base = np.array([1.7, -1.5, 0.3])[:bathnum]
    return base[:, None] + 0.2 * rng.standard_normal((bathnum, nruns)), T


def write_kappa_files(vals, T):
    for i in range(vals.shape[0]):
        for j in range(vals.shape[1]):
            with open("kappa." + str(T) + ".bath" + str(i) + ".run" + str(j) + ".dat", "w") as f:
                f.write("%i %f    %f \n" % (j, T, vals[i, j]))


def phbath_sig_inputs():
    """a seeded retarded self-energy on a frequency grid that starts at 0 (ggamma copies index 1 there, baths.py:386-388):
    Sigma(w) = Re - i w Gamma(w) with Gamma PSD, in the MD units (eV^2 / eV)"""
    nc, ngw = 5, 9
    gwl, g = gamma_grid(ngw, nc, 91, wmax=0.24)
    re = np.array([sym(nc, 92 + i, 0.01) for i in range(ngw)])
    sig = re - 1j * gwl[:, None, None] * g
    return nc, gwl, sig


def write_classic_nc(filename, variables):
    """every array as a float64 (or int32) variable of a NetCDF classic file, one dimension per axis"""
    from scipy.io import netcdf_file
    f = netcdf_file(filename, 'w')
    for name, arr in variables.items():
        arr = np.asarray(arr)
        code = 'i' if arr.dtype.kind in 'iu' else 'd'
        arr = arr.astype(np.int32 if code == 'i' else float)
        dims = []
        for ax, ln in enumerate(arr.shape):
            d = "%s_d%d" % (name, ax)
            f.createDimension(d, ln)
            dims.append(d)
        v = f.createVariable(name, code, tuple(dims))
        v[:] = arr
    f.close()
