"""Host-side functions of the PRODUCT package (sclmd_b200) pinned to outputs of the reference itself (tests/golden/*.npz,
written by oracle/make_golden.py): the scalar helpers of the noise spectrum, the post-processing of the kappa.* files
(tools.calHF / calTC, tools.py:132-215) and gamma(w) from a lead self-energy (phbath.ggamma, baths.py:375-395).
None of these needs a device."""
import os

import numpy as np

import problems as P


def test_scalar_helpers_of_the_product_equal_the_reference(golden_dir):
    from sclmd_b200 import functions as F
    from sclmd_b200 import noise as N
    g = np.load(os.path.join(golden_dir, "scalars.npz"))
    for w, T, cl, zp, r in g["equ"]:                       # noise.equ (noise.py:249-270) incl. the bose edge cases it calls
        with np.errstate(all="ignore"):
            o = N.equ(w, 0.4, T, bool(cl), bool(zp))
        assert o == r or (np.isnan(o) and np.isnan(r)), (w, T, cl, zp)
    assert np.array_equal(np.array([F.flinterp(x, g["xs"], g["ys"]) for x in g["xq"]]), g["fl"])      # functions.py:117-134
    assert np.array_equal(np.array([F.nearest(x, g["xs"]) for x in g["xq"]]), g["nn"])                # functions.py:137-143
    for x, v in zip(g["xq"], g["fl"]):                    # the index form the device plans are built from
        i0, i1, wt = F.flinterp_index(x, g["xs"])
        assert abs(g["ys"][i0] + wt * (g["ys"][i0] - g["ys"][i1]) - v) < 1e-15
    # functions.bose (functions.py:80-99): by-fiat values
    assert F.bose(0.0, 300.0) == 0.0 and F.bose(-0.1, 0.0) == -1.0 and F.bose(0.1, 0.0) == 0.0


def test_calHF_and_calTC_equal_the_reference(golden_dir, tmp_path, monkeypatch):
    from sclmd_b200 import tools
    g = np.load(os.path.join(golden_dir, "tools.npz"))
    for name, seed, nb, nr, T, kw in P.KAPPA_CASES:
        d = tmp_path / name
        d.mkdir()
        monkeypatch.chdir(d)
        vals, T = P.kappa_case(seed, nb, nr, T)
        P.write_kappa_files(vals, T)
        tools.calHF(dlist=kw["dlist"], bathnum=nb)
        tools.calTC(kw["delta"], dlist=kw["dlist"], bathnum=nb, L=kw.get("L"), A=kw.get("A"))
        assert np.array_equal(np.loadtxt("heatflux.%d.dat" % T), g[name + "_hf"]), name
        assert np.array_equal(np.loadtxt("heatflux-between-baths.%d.dat" % T), g[name + "_jb"]), name
        if kw["delta"] != 0:
            assert np.array_equal(np.loadtxt("thermalconductance.%d.dat" % T), g[name + "_tc"]), name
        else:
            assert not os.path.exists("thermalconductance.%d.dat" % T)
        if kw.get("L") is not None:
            assert np.array_equal(np.loadtxt("thermalconductivity.%d.dat" % T), g[name + "_cond"]), name


def test_ggamma_equals_the_reference(golden_dir):
    from sclmd_b200.baths import phbath
    g = np.load(os.path.join(golden_dir, "phbath_sig.npz"))
    nc, gwl, sig = P.phbath_sig_inputs()
    b = phbath(300.0, list(range(nc)), 0.06, 48, 0.25 / 0.658, 64, ml=11, mcof=2.0, sig=sig, gwl=gwl)
    assert np.array_equal(b.gamma, g["gamma_eta0"])
    assert np.array_equal(b.gamma[0], b.gamma[1])          # the w == 0 node is a copy of the next one (baths.py:386-388)


def test_c4_lambda_fixture_is_pinned_to_raw_file_offsets(golden_dir):
    """make_golden.py locates the five matrices in the example's NetCDF-4 file by a raw byte scan that parses no HDF5 and
    requires them to equal the fixture bit for bit, in file order; the offsets are recorded"""
    o = np.load(os.path.join(golden_dir, "c4_lambda_offsets.npz"))
    order = sorted(o.files, key=lambda k: int(o[k]))
    assert order == ["eta_r", "xim_r", "xip_r", "zeta1_r", "zeta2_r"]
    assert all(int(o[b]) - int(o[a]) >= 36 * 36 * 8 for a, b in zip(order, order[1:]))
