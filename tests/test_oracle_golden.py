"""The CPU oracle (oracle/sclmd_oracle.py) replayed against fixtures written by the
REFERENCE itself (oracle/make_golden.py; reference imported in place from
/root/reference in the build container).  CPU only."""
import os

import numpy as np
import pytest

import problems as P
from oracle import sclmd_oracle as O


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def oracle_baths(c):
    """bath inputs as the reference holds them after CheckEmat / gmem."""
    out = []
    for b in range(len(c["cids"])):
        if c["kinds"][b] == "ph":
            out.append(dict(kind="ph", cids=c["cids"][b], kernel=c["kern"][b], bias=0.0, exim=None, zeta1=None, zeta2=None))
        else:
            e = c["e"]
            sym = lambda a: None if a is None else 0.5 * (a + a.T)
            asym = lambda a: None if a is None else 0.5 * (a - a.T)
            out.append(dict(kind="e", cids=c["cids"][b], kernel=np.array([sym(e["efric"][b])]), bias=e["bias"][b],
                            exim=asym(e["exim"][b]), zeta1=sym(e["zeta1"][b]), zeta2=asym(e["zeta2"][b])))
    return out


@pytest.mark.parametrize("name", list(P.MD_CASES))
def test_md_literal_and_ensemble_match_reference(name, golden_dir):
    c = P.MD_CASES[name]()
    g = np.load(os.path.join(golden_dir, "md_%s.npz" % name))
    K = P.psd_project(c["K"])
    assert abs(float(np.sum(K * K)) - float(g["dyn_checksum"])) <= 1e-12 * float(g["dyn_checksum"])
    ob = oracle_baths(c)
    lit = O.LiteralMD(K, c["dt"], c["nmd"],
                      [O.Bath(o["kind"], o["cids"], o["kernel"], c["noise"][i], c["dt"], c["nmd"], o["bias"],
                              o["exim"], o["zeta1"], o["zeta2"]) for i, o in enumerate(ob)], c["cons"])
    ntraj = 3
    ens = O.EnsembleMD(K, c["dt"], c["nmd"], ntraj, c["cons"])
    for i, o in enumerate(ob):
        # trajectory 1 carries the golden noise, the others something else (independence check)
        nz = np.stack([0.5 * c["noise"][i], c["noise"][i], -c["noise"][i]])
        ens.add_bath(o["cids"], o["kernel"], nz, o["bias"], o["exim"], o["zeta1"], o["zeta2"], o["kind"])
    lit.q, lit.p = g["q0"].copy(), g["p0"].copy()
    ens.q[:], ens.p[:] = g["q0"], g["p0"]
    n = int(g["nsteps"])
    full = g["q"].shape[0] == n
    for s in range(n):
        lit.vv()
        ens.step()
        if full or s == n - 1:
            k = s if full else 0
            assert relerr(lit.q, g["q"][k]) < 1e-11 and relerr(lit.p, g["p"][k]) < 1e-11
            assert relerr(ens.q[1], g["q"][k]) < 1e-11 and relerr(ens.p[1], g["p"][k]) < 1e-11
    assert relerr(np.array([b.cur for b in lit.baths]), g["cur"]) < 1e-10
    assert relerr(np.array([b["cur"][1] for b in ens.baths]), g["cur"]) < 1e-10
    assert relerr(ens.etot[1], g["etot"]) < 1e-11
    assert lit.t == n == ens.t


def test_diag_kernel_equals_full_diagonal():
    """diagonal-kernel storage is the full kernel with a diagonal matrix."""
    c = P.md_case_ph_full()
    K = P.psd_project(c["K"])
    a = O.EnsembleMD(K, c["dt"], c["nmd"], 2, c["cons"])
    b = O.EnsembleMD(K, c["dt"], c["nmd"], 2, c["cons"])
    for i in range(2):
        kd = P.diag_kernel(7, 6, c["dt"], seed=3 + i)
        kf = np.array([np.diag(r) for r in kd])
        nz = P.injected_noise(2, c["nmd"], 6, seed=9 + i)
        a.add_bath(c["cids"][i], kd, nz)
        b.add_bath(c["cids"][i], kf, nz)
    a.run(45)
    b.run(45)
    assert relerr(a.q, b.q) < 1e-13 and relerr(a.p, b.p) < 1e-13


def test_noise_replay(golden_dir):
    g = np.load(os.path.join(golden_dir, "noise.npz"))
    dt, nmd = 0.25 / 0.658, 32
    z = np.random.default_rng(60).standard_normal(4096)

    def replay(av, au, zz):
        k = [0]

        def draw(scale):
            v = scale * zz[k[0]]
            k[0] += 1
            return v
        x = [O.vargau(av[i], au[i], draw) for i in range(nmd // 2 + 1)]
        return O.spectrum_to_series(np.array(x), dt, nmd), k[0]
    ph, used = replay(g["ph_av"], g["ph_au"], z)
    assert used == int(g["used_ph"]) and relerr(ph, g["ph"]) < 1e-13
    en, used = replay(g["e_av"], g["e_au"], z[1000:])
    assert used == int(g["used_e"]) and relerr(en, g["en"]) < 1e-13
    # covariance restatement == the reference's captured eigensystems
    gwl, gam = P.gamma_grid(7, 4, 61, wmax=0.3)
    for i in range(nmd // 2 + 1):
        A = O.ph_covariance(i, gam, gwl, 300.0, float(g["phcut"]), dt, nmd)
        assert np.max(np.abs((g["ph_au"][i] * g["ph_av"][i]) @ g["ph_au"][i].conj().T - A)) < 1e-12
    efric, exim, exip = P.psd(3, 62, 0.05), P.antisym(3, 63, 0.01), P.sym(3, 64, 0.01)
    for i in range(nmd // 2 + 1):
        A = O.e_covariance(i, efric, exim, exip, 0.2, 300.0, 2.0, dt, nmd, False, False)
        assert np.max(np.abs((g["e_au"][i] * g["e_av"][i]) @ g["e_au"][i].conj().T - A)) < 1e-12
    # the factor form used by the CUDA path gives the same series as vargau with indexed draws
    L = np.array([O.eig_factor(O.ph_covariance(i, gam, gwl, 300.0, float(g["phcut"]), dt, nmd))[0] for i in range(nmd // 2 + 1)])
    xi = z[:(nmd // 2 + 1) * 4].reshape(nmd // 2 + 1, 4)
    s = O.noise_from_factors(L, xi, dt, nmd)
    assert np.all(np.isfinite(s)) and relerr(s, s[(-np.arange(nmd)) % nmd]) < 1e-12   # real spectrum -> even series


def test_scalars(golden_dir):
    g = np.load(os.path.join(golden_dir, "scalars.npz"))
    for w, T, cl, zp, r in g["equ"]:
        with np.errstate(all="ignore"):
            o = O.equ(w, 0.4, T, bool(cl), bool(zp))
        assert o == r or (np.isnan(o) and np.isnan(r))
    assert np.array_equal(np.array([O.flinterp(x, g["xs"], g["ys"]) for x in g["xq"]]), g["fl"])
    assert np.array_equal(np.array([O.nearest(x, g["xs"]) for x in g["xq"]]), g["nn"])
    gwl, gam = P.gamma_grid(6, 3, 70, wmax=0.25)
    wl = [0.3 * i / 40 for i in range(40)]
    tl = [0.38 * i for i in range(9)]
    assert relerr(O.gamt(tl, wl, gwl, gam, 0), g["gamt0"]) < 1e-13
    assert relerr(O.gamt(tl, wl, gwl, gam, 0.01), g["gamt1"]) < 1e-13
    assert O.equ(0.0, 1.0, 300.0) == 2 * O.KB * 300.0


def test_bpt(golden_dir):
    g = np.load(os.path.join(golden_dir, "bpt.npz"))
    K = P.spring_chain_dyn(12, seed=80) / O.RPC ** 2
    fixed = list(range(0, 3)) + list(range(33, 36))
    Kr = np.delete(np.delete(0.5 * (K + K.T), fixed, 0), fixed, 1)
    iL, iR = O.bpt_reduce_index(range(3, 12), 3), O.bpt_reduce_index(range(24, 33), 3)
    tm = np.array([O.bpt_tm(Kr, w, 0.1, iL, iR) for w in g["tm"][:, 0]])
    assert relerr(tm, g["tm"][:, 1]) < 1e-9
    assert np.array_equal(g["tm"][:, 0], np.linspace(0, 0.25 / O.RPC, 21))
    ps = np.array([O.bpt_ps_nobias(Kr, w, 300.0, 0.1, iL, iR, np.arange(30)) for w in g["ps"][:, 0]])
    assert relerr(ps[1:], g["ps"][1:, 1]) < 1e-9
    kap = np.array([O.thermalcurrent(g["tm"], T, 0.1) / (T * 0.1) for T in (100.0, 300.0, 900.0)])
    assert relerr(kap, g["kappa"]) < 1e-12
    # biased electron bath on the centre block (negf.py:162-193, 234-236)
    bd, cp, cm = P.psd(6, 82, 2.0), P.sym(6, 83, 1.5), P.antisym(6, 84, 1.5)
    with np.errstate(all="ignore"):
        psb = np.array([O.bpt_ps_bias(Kr, w, 300.0, 0.1, iL, iR, 12, bd, cp, cm, 0.6 / O.RPC, np.arange(12, 18)) for w in g["ps_bias"][:, 0]])
        tmb = np.array([O.bpt_tm_bias(Kr, w, 300.0, 0.1, iL, iR, 12, bd, cp, cm, 0.6 / O.RPC) for w in g["tm_bias"][:, 0]])
    assert relerr(psb[1:], g["ps_bias"][1:, 1]) < 1e-9 and relerr(tmb, g["tm_bias"][:, 1]) < 1e-9


def test_sig(golden_dir):
    g = np.load(os.path.join(golden_dir, "sig.npz"))
    K00, K11, K01 = P.chain_blocks(4, seed=81)
    K10 = K01.T
    eta = float(g["eta"])
    for d, key in (("L", "seL"), ("R", "seR")):
        se = np.array([O.sig_selfenergy(K00, K11, K01, K10, w, eta, d) for w in g["ep"]])
        assert relerr(se, g[key]) < 1e-10
    tm = np.array([O.sig_tm(K00, K11, K01, K10, w, eta) for w in g["ep"]])
    assert relerr(tm, g["tm"][:, 1]) < 1e-8
    # physics KAT: a perfect chain transmits one channel per dof inside the band
    assert 3.9 < tm[3] < 4.0001


def test_noise_replay_config4_matrices(golden_dir):
    """enoise of the biased junction bath of the current-induced example (rundp.py:76-77) with its own 36x36 matrices:
    the oracle rebuilds the reference's series from the captured eigensystems and its covariance restatement matches"""
    g = np.load(os.path.join(golden_dir, "noise_c4.npz"))
    lam = P.c4_lambda()
    dt, nmd = 0.5 / 0.658, 16
    z = np.random.default_rng(66).standard_normal(4096)
    k = [0]

    def draw(scale):
        v = scale * z[k[0]]
        k[0] += 1
        return v
    x = [O.vargau(g["e_av"][i], g["e_au"][i], draw) for i in range(nmd // 2 + 1)]
    en = O.spectrum_to_series(np.array(x), dt, nmd)
    assert k[0] == int(g["used"]) and relerr(en, g["en"]) < 1e-12
    for i in range(nmd // 2 + 1):
        A = O.e_covariance(i, lam["eta_r"], lam["xim_r"], lam["xip_r"], 1.0, 300.0, 2.0, dt, nmd, False, False)
        assert np.max(np.abs((g["e_au"][i] * g["e_av"][i]) @ g["e_au"][i].conj().T - A)) < 1e-18


def test_netcdf4_reader_against_fixture(golden_dir):
    """sclmd_b200.myio reads the reference's NetCDF-4 (HDF5) example file without netCDF4 / h5py; the fixture the GPU box
    uses is what it read.  Needs the reference tree (this container only)."""
    path = "/root/reference/examples/current-induced/grapheneLambda-r-0.3-ver2.nc"
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    from sclmd_b200 import myio
    v = myio.read_nc_variables(path)
    lam = P.c4_lambda()
    for k in ("eta_r", "xim_r", "xip_r", "zeta1_r", "zeta2_r"):
        assert v[k].shape == (36, 36) and np.array_equal(v[k], lam[k])
    assert v["hw"].shape == (36,) and v["U"].shape == (36, 36) and v["blist"].shape == (400,)
    assert abs(np.abs(v["U"] @ v["U"].T - np.eye(36)).max()) < 1e-10           # the mode matrix it ships is orthogonal
    ds = myio.Dataset(path, "r")                                                 # the access pattern of rundp.py:10,76
    assert np.array_equal(ds["eta_r"][:], lam["eta_r"])
    with pytest.raises(ValueError):
        myio.read_nc_variables(__file__)


def test_eph_file_round_trip(tmp_path):
    """myio.WriteEPHNCfile / ReadNewEPHNCFile (myio.py:109-171): NetCDF classic on the way out, either flavour on the way in"""
    from sclmd_b200 import myio
    rng = np.random.default_rng(8)
    n, nw, ns = 6, 5, 3
    args = dict(wl=np.linspace(0, 0.2, nw), hw=rng.random(n), U=rng.standard_normal((n, n)), DynMat=P.sym(n, 1),
                SigL=rng.standard_normal((nw, ns, ns)) + 1j * rng.standard_normal((nw, ns, ns)),
                SigR=rng.standard_normal((nw, ns, ns)) + 1j * rng.standard_normal((nw, ns, ns)),
                Friction=P.psd(n, 2), NC=P.antisym(n, 3), NCP=P.sym(n, 4), zeta1=P.sym(n, 5), zeta2=P.antisym(n, 6))
    fn = str(tmp_path / "eph.nc")
    myio.WriteEPHNCfile(fn, **args)
    e = myio.ReadNewEPHNCFile(fn)
    assert np.array_equal(e.wl, args["wl"]) and np.array_equal(e.DynMat, args["DynMat"]) and np.array_equal(e.SigL, args["SigL"])
    assert np.array_equal(e.efric, args["Friction"]) and np.array_equal(e.xim, args["NC"]) and np.array_equal(e.zeta2, args["zeta2"])
    assert np.array_equal(myio.ReadNetCDFVar(fn, "hw"), args["hw"])


def test_myio_readers_match_the_reference(tmp_path, golden_dir):
    """ReadLambda / ReadwbLambda / ReadDynmat / ReadSig / ord2idx (myio.py:214-366) against the reference's own outputs on the
    same seeded files (oracle/make_golden.py:run_myio_cases)"""
    from sclmd_b200 import myio
    g = np.load(os.path.join(golden_dir, "myio_readers.npz"))
    inp = P.myio_inputs()
    for kind in ("lam", "wb", "ph", "sg"):
        P.write_classic_nc(str(tmp_path / (kind + ".nc")), inp[kind])
    for w0 in (0.0, 0.07, 0.5):
        r = myio.ReadLambda(str(tmp_path / "lam.nc"), w0)
        want = g["lam_%g" % w0]
        assert r[0] == want[0][0, 0]
        for k in range(5):
            assert np.array_equal(r[1 + k], want[1 + k])
    r = myio.ReadwbLambda(str(tmp_path / "wb.nc"))
    assert r[0] == 0.0 and all(np.array_equal(r[1 + k], g["wb"][1 + k]) for k in range(5))
    for tag, order in (("plain", None), ("reordered", [2, 1])):
        dyn, U, hw = myio.ReadDynmat(str(tmp_path / "ph.nc"), order)
        assert np.array_equal(U, g["U_" + tag]) and np.array_equal(hw, g["hw_" + tag])
        assert np.max(np.abs(dyn - g["dyn_" + tag])) <= 1e-14 * np.max(np.abs(g["dyn_" + tag]))
    e = myio.ReadSig(str(tmp_path / "sg.nc"))
    assert np.array_equal(e.wl, g["sig_wl"]) and np.array_equal(e.SigL, g["sigL"]) and np.array_equal(e.SigR, g["sigR"])
    assert np.array_equal(myio.ord2idx([3, 1, 2]), g["ord2idx"])
    with pytest.raises(ValueError):
        myio.ReadDynmat(str(tmp_path / "ph.nc"), [1, 2, 3])


def test_small_tools(tmp_path, monkeypatch):
    """tools.get_atomname / get_atommass / eff / avdf (tools.py:7-32, 218-259) against their definitions"""
    from sclmd_b200 import tools as T
    monkeypatch.chdir(tmp_path)
    assert T.get_atommass("C") == 12.0107 and T.get_atomname(12.011) == "C" and T.get_atomname(196.97) == "Au"
    assert T.get_atommass("Xx") is None and T.get_atomname(500.0) is None
    rng = np.random.default_rng(4)
    a = rng.standard_normal((6, 6))
    dyn = a @ a.T - 0.8 * np.eye(6)                              # one or two negative eigenvalues
    assert (np.linalg.eigvalsh(dyn) < 0).any()
    np.savetxt("dynmat.dat", (dyn + 1e-3 * rng.standard_normal((6, 6))).reshape(-1, 3))
    out = T.eff("dynmat.dat")
    w = np.linalg.eigvalsh(out)
    assert np.allclose(out, out.T) and w.min() > -1e-12 and np.allclose(np.loadtxt("moddynmat.dat"), out)
    f0, f1 = rng.standard_normal((5, 4)), rng.standard_normal((5, 4))
    np.save("deltaforce.run0.npy", f0)
    np.save("deltaforce.run1.npy", f1)
    T.avdf(["deltaforce.run0.npy", "deltaforce.run1.npy"], outputname="df", abs=True)
    both = np.abs(np.concatenate((f0, f1)))
    assert np.allclose(np.loadtxt("df-mean0.dat"), np.abs(f0).mean(axis=0))
    assert np.allclose(np.loadtxt("df-mean1.dat"), both.mean(axis=0))
    assert np.allclose(np.loadtxt("df-deviation1.dat"), both.std(axis=0))


def test_bpt_and_sig_write_the_reference_frequency_files(tmp_path, monkeypatch):
    """negf.py:92-102 / selfenergy.py:76-91: omegas.dat (signed sqrt of the eigenvalues times hbar), eigvecs.dat and
    falsefrequencies.dat appear when the dynamical matrix is loaded from a file (host code, no device needed)"""
    from sclmd_b200.negf import bpt
    from sclmd_b200.selfenergy import sig
    monkeypatch.chdir(tmp_path)
    K = P.spring_chain_dyn(12, seed=80)
    K[5, 5] -= 3.0 * abs(K[5, 5])                                              # force a false (imaginary) frequency
    K = 0.5 * (K + K.T)
    np.savetxt("dyn.dat", K.reshape(-1, 3))
    fixed = [list(range(0, 3)), list(range(33, 36))]
    b = bpt(None, 0.25, 0.1, [list(range(3, 12)), list(range(24, 33))], fixed, dynmatfile="dyn.dat", num=4)
    red = np.delete(np.delete(K, fixed[0] + fixed[1], axis=0), fixed[0] + fixed[1], axis=1)
    w = np.linalg.eigvalsh(red)
    want = np.where(w > 0, np.sqrt(np.abs(w)), -np.sqrt(np.abs(w))) * b.rpc
    assert np.allclose(np.loadtxt("omegas.dat"), want) and np.allclose(b.omegas, want)
    assert np.loadtxt("eigvecs.dat").shape == (30, 30)
    assert list(np.atleast_1d(np.loadtxt("falsefrequencies.dat", dtype=int))) == list(np.nonzero(~(w > 0))[0])
    for f in ("omegas.dat", "eigvecs.dat", "falsefrequencies.dat"):
        os.remove(f)
    s = sig(None, 0.06, range(3, 12), range(12, 21), dofatomfixed=fixed, dynmatfile="dyn.dat", num=4, eta=1e-3)
    assert np.allclose(np.loadtxt("omegas.dat"), want * s.rpc / b.rpc) and os.path.exists("eigvecs.dat")
    os.remove("omegas.dat")
    bpt(None, 0.25, 0.1, [list(range(3, 12)), list(range(24, 33))], fixed, dynmatfile=K, num=4)   # array in: nothing written
    assert not os.path.exists("omegas.dat")


def test_reordxyz():
    """myio.reordxyz (myio.py:64-77)"""
    from sclmd_b200 import myio
    anr = [6, 6, 1, 1, 79, 79]
    xyz = [[float(i), 0.0, 0.0] for i in range(6)]
    a2, x2 = myio.reordxyz(anr, xyz, [4, 2, 3])
    assert a2 == [6, 1, 6, 1, 79, 79] and [x[0] for x in x2] == [0.0, 3.0, 1.0, 2.0, 4.0, 5.0]
    with pytest.raises(ValueError):
        myio.reordxyz(anr, xyz, [2, 5])


def test_modal_restatement_equals_the_real_space_one():
    """oracle.ModalMD (propagation in the eigenbasis of md.setDyn, the specification of the engine's modal mode) against
    oracle.EnsembleMD over 1024 steps: the difference stays at the level a 1-ulp change of K produces"""
    natoms, nc, ml, ntraj, nmd = 30, 10, 40, 2, 64
    nph, dt = 3 * natoms, 0.25 / 0.658
    Kraw = P.spring_chain_dyn(natoms, seed=5)
    lam, U = np.linalg.eigh(0.5 * (Kraw + Kraw.T))
    lam = np.where(lam < 0, 0.0, lam)
    K = U @ np.diag(lam) @ U.T
    cids = [list(range(0, nc)), list(range(nph - nc, nph))]
    ens, mod = O.EnsembleMD(K, dt, nmd, ntraj, None), O.ModalMD(lam, U, dt, nmd, ntraj)
    for b in range(2):
        kern, nz = P.diag_kernel(ml if b == 0 else 1, nc, dt, 50 + b), P.injected_noise(ntraj, nmd, nc, seed=60 + b)
        ens.add_bath(cids[b], kern, nz)
        mod.add_bath(cids[b], kern, nz)
    rng = np.random.default_rng(7)
    q0, p0 = 0.05 * rng.standard_normal((ntraj, nph)), 0.02 * rng.standard_normal((ntraj, nph))
    ens.q[:], ens.p[:] = q0, p0
    mod.set_state(q0, p0)
    for s in range(1024):
        ens.step()
        mod.step()
        if s in (0, 1, 63, 1023):
            q, p = mod.state()
            assert relerr(q, ens.q) < 1e-12 and relerr(p, ens.p) < 1e-12, s
    keep = [k for k in range(nmd) if k != 1024 % nmd]          # ModalMD records evaluation A of the NEXT step at the end of a step
    assert relerr(mod.etot[:, keep], ens.etot[:, keep]) < 1e-11
    for b in range(2):
        assert relerr(mod.baths[b]["cur"][:, keep], ens.baths[b]["cur"][:, keep]) < 1e-11


def test_time_blocked_tail_decomposition_equals_the_direct_tail():
    """near + mid + far-far (in any split into age ranges) is the direct tail of md.py:386-387 for every step of a block, also when the
    ring has wrapped and slots of the oldest rows have been overwritten by the block's own rows; and the Hankel-fragment identity the
    tensor-pipe far pass relies on: the tile of step tile n at ages d0..d0+3 is the fragment F(d0 + 8 n)"""
    rng = np.random.default_rng(5)
    ml, nc, dt, tb = 96, 3, 0.37, 32
    kern = rng.standard_normal((ml, nc)) * np.exp(-np.arange(ml) / 40.0)[:, None]
    for t0 in (0, 32, 96, 160):
        p_of = {tt: rng.standard_normal(nc) for tt in range(t0 - 2 * ml, t0 + tb)}
        for s in (0, 1, 7, 31):
            t = t0 + s
            ring = np.zeros((ml, nc))
            for tt in range(t - ml + 1, t + 1):                # the ring holds the last ml rows
                ring[tt % ml] = p_of[tt]
            direct = dt * sum(kern[j] * p_of[t + 1 - j] for j in range(1, ml))
            near, mid, far = O.blocked_tail_parts(kern, ring, t0, s, dt, tb)
            assert np.max(np.abs(near + mid + far - direct)) <= 1e-13 * np.max(np.abs(direct))
            _, _, far3 = O.blocked_tail_parts(kern, ring, t0, s, dt, tb, seg=[(32, 48), (48, 80), (80, ml)])
            assert np.max(np.abs(far3 - far)) <= 1e-13 * max(np.max(np.abs(far)), 1e-300)
    # Hankel fragments: H[s_global][d] = k[s_global + 2 + d]; tile (n, d0) == F(d0 + 8 n)
    kvec = rng.standard_normal(200)
    for n in range(4):
        for d0 in (0, 4, 40):
            tile = np.array([[kvec[(8 * n + s) + 2 + (d0 + a)] for a in range(4)] for s in range(8)])
            assert np.array_equal(tile, O.hankel_tile(kvec, d0 + 8 * n, n))
