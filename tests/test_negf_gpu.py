"""NEGF transmission / power spectrum and surface self-energy sweeps on the device vs the
reference's golden numbers and the oracle."""
import os

import numpy as np
import pytest

import problems as P
from oracle import sclmd_oracle as O

pytestmark = pytest.mark.gpu


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def test_bpt_golden(golden_dir, tmp_path, monkeypatch):
    from sclmd_b200.negf import bpt
    monkeypatch.chdir(tmp_path)
    g = np.load(os.path.join(golden_dir, "bpt.npz"))
    K = P.spring_chain_dyn(12, seed=80) / O.RPC ** 2
    fixed = [list(range(0, 3)), list(range(33, 36))]
    bath = [list(range(3, 12)), list(range(24, 33))]
    b = bpt(None, 0.25, 0.1, bath, fixed, dynmatfile=K, num=20)
    b.gettm()
    assert np.array_equal(b.tmnumber[:, 0], g["tm"][:, 0])          # frequency grid bit-exact
    assert relerr(b.tmnumber[:, 1], g["tm"][:, 1]) < 1e-8
    assert os.path.exists("transmission.dat")
    b.getps(300.0, 0.25, 20)
    assert relerr(b.psnumber[1:, 1], g["ps"][1:, 1]) < 1e-8
    kap = np.array([b.thermalconductance(T, 0.1) for T in (100.0, 300.0, 900.0)])
    assert relerr(kap, g["kappa"]) < 1e-8
    assert abs(b.tm(0.0)) < 1e-20                                      # Gamma ~ w -> T(0) = 0


def test_bpt_with_biased_electron_bath(golden_dir, tmp_path, monkeypatch):
    """bpt.setbias + getps/gettm (examples/current-induced/runnegf.py flow) against the reference's golden numbers"""
    from sclmd_b200.negf import bpt
    monkeypatch.chdir(tmp_path)
    g = np.load(os.path.join(golden_dir, "bpt.npz"))
    K = P.spring_chain_dyn(12, seed=80) / O.RPC ** 2
    fixed = [list(range(0, 3)), list(range(33, 36))]
    bath = [list(range(3, 12)), list(range(24, 33))]
    b = bpt(None, 0.25, 0.1, bath, fixed, dynmatfile=K, num=20)
    bd, cp, cm = P.psd(6, 82, 2.0), P.sym(6, 83, 1.5), P.antisym(6, 84, 1.5)
    with pytest.raises(ValueError):
        b.setbias(0.6, bdamp=bd, chiplus=cp, chiminus=cm, dofatomofbias=list(range(15, 20)))
    b.setbias(0.6, bdamp=bd, chiplus=cp, chiminus=cm, dofatomofbias=list(range(15, 21)))
    b.getps(300.0, 0.25, 20, atomlist=list(range(15, 21)), filename="bias")
    assert relerr(b.psnumber[1:, 1], g["ps_bias"][1:, 1]) < 1e-8
    b.gettm()
    assert relerr(b.tmnumber[:, 1], g["tm_bias"][:, 1]) < 1e-8
    # zero bias keeps the block damping but no nonequilibrium terms
    b.setbias(0.0, bdamp=bd, chiplus=cp, chiminus=cm, dofatomofbias=list(range(15, 21)))
    om = np.array([17.0, 88.0, 250.0])
    got = b.ps_sweep(om, 300.0, list(range(15, 21)))
    iL, iR = O.bpt_reduce_index(bath[0], 3), O.bpt_reduce_index(bath[1], 3)
    want = np.array([O.bpt_ps_bias(b.dynmat, w, 300.0, 0.1, iL, iR, 12, bd, cp, cm, 0.0, np.arange(12, 18)) for w in om])
    assert relerr(got, want) < 1e-8


def test_bpt_config3_shape_vs_oracle():
    """n = 483 (examples/runnegf.py shape): Gamma on 150 + 150 dofs, several panels, pivoting"""
    from sclmd_b200.negf import bpt
    natoms = 201
    K = P.spring_chain_dyn(natoms, seed=14) / O.RPC ** 2
    fixed = [list(range(0, 60)), list(range(543, 603))]
    bath = [list(range(60, 210)), list(range(393, 543))]
    b = bpt(None, 0.25, 0.1, bath, fixed, dynmatfile=K, num=1000)
    om = np.array([0.0, 3.7, 41.3, 97.0, 150.2, 233.3, 301.9])
    got = b.tm_sweep(om)
    iL, iR = O.bpt_reduce_index(bath[0], 60), O.bpt_reduce_index(bath[1], 60)
    want = np.array([O.bpt_tm(b.dynmat, w, 0.1, iL, iR) for w in om])
    assert np.max(np.abs(got - want)) < 1e-8 * max(1.0, np.abs(want).max())
    sel = list(range(60 + 150, 60 + 333))
    ps = b.ps_sweep(om[1:4], 300.0, sel)
    wantps = np.array([O.bpt_ps_nobias(b.dynmat, w, 300.0, 0.1, iL, iR, np.array(sel) - 60) for w in om[1:4]])
    assert relerr(ps, wantps) < 1e-8


def test_bpt_large_system_fallback_panel():
    """n = 654 (config-4 size) exceeds the register-resident panel (n <= 512): shared-memory panel path"""
    from sclmd_b200.negf import bpt
    natoms = 242
    K = P.spring_chain_dyn(natoms, seed=15) / O.RPC ** 2
    fixed = [list(range(0, 24)), list(range(678, 726))]
    bath = [list(range(24, 144)), list(range(558, 678))]
    b = bpt(None, 0.25, 0.1, bath, fixed, dynmatfile=K, num=10)
    assert len(b.dynmat) == 654
    om = np.array([5.0, 120.7, 290.1])
    got = b.tm_sweep(om)
    iL, iR = O.bpt_reduce_index(bath[0], 24), O.bpt_reduce_index(bath[1], 24)
    want = np.array([O.bpt_tm(b.dynmat, w, 0.1, iL, iR) for w in om])
    assert np.max(np.abs(got - want)) < 1e-8 * max(1.0, np.abs(want).max())


def test_sig_golden(golden_dir, tmp_path, monkeypatch):
    from sclmd_b200.selfenergy import sig
    monkeypatch.chdir(tmp_path)
    g = np.load(os.path.join(golden_dir, "sig.npz"))
    K00, K11, K01 = P.chain_blocks(4, seed=81)
    full = np.zeros((8, 8))
    full[:4, :4], full[4:, 4:], full[:4, 4:], full[4:, :4] = K00, K11, K01, K01.T
    s = sig(None, 0.06, range(0, 4), range(4, 8), dynmatfile=full, num=8, eta=2e-3)
    assert np.array_equal(s.ep, g["ep"])
    seL = s.getse('L')
    assert relerr(s.dos, g["dosL"]) < 1e-9
    seR = s.getse('R')
    assert relerr(seL, g["seL"]) < 1e-10 and relerr(seR, g["seR"]) < 1e-10
    s.gettm()
    assert relerr(s.tmnumber[:, 1], g["tm"][:, 1]) < 1e-8
    its = np.array([O.sig_sgf(K00, K11, s.K01, s.K10, w, s.eta, "R")[1] for w in s.ep])
    assert np.array_equal(s.iterations, its)                            # same iteration counts as the reference loop


def test_sig_24x24_vs_oracle():
    from sclmd_b200.selfenergy import sig
    m = 24
    K00, K11, K01 = P.chain_blocks(m, seed=5, k2=0.08)
    full = np.zeros((2 * m, 2 * m))
    full[:m, :m], full[m:, m:], full[:m, m:], full[m:, :m] = K00, K11, K01, K01.T
    s = sig(None, 0.06, range(0, m), range(m, 2 * m), dynmatfile=full, num=40, eta=1e-3)
    om = s.ep[[0, 3, 11, 25, 40]]
    se = s.selfenergy_sweep(om, 'R')
    want = np.array([O.sig_selfenergy(K00, K11, s.K01, s.K10, w, s.eta, "R") for w in om])
    assert relerr(se, want) < 1e-9
    tm = s.tm_sweep(om)
    wtm = np.array([O.sig_tm(K00, K11, s.K01, s.K10, w, s.eta) for w in om])
    assert np.max(np.abs(tm - wtm)) < 1e-8 * max(1.0, np.abs(wtm).max())
    assert np.array_equal(s.iterations, np.array([O.sig_sgf(K00, K11, s.K01, s.K10, w, s.eta, 'R')[1] for w in om]))
    # sig.sgf (selfenergy.py:105-131): the surface Green function itself, both leads
    for d in ('L', 'R'):
        gs = s.sgf(float(om[3]), d)
        wg, wit = O.sig_sgf(K00, K11, s.K01, s.K10, float(om[3]), s.eta, d)
        assert relerr(gs, wg) < 1e-9 and int(s.iterations[0]) == wit
    with pytest.raises(ValueError):
        s.sgf(float(om[3]), 'X')
    # sig.retargf (selfenergy.py:145-147) from the device and the reference's own tm formula built on it (selfenergy.py:149-151)
    w = float(om[2])
    sl, sr = s.selfenergy(w, 'L'), s.selfenergy(w, 'R')
    g = s.retargf(w)
    assert relerr(g, np.linalg.inv((w + 1e-8j) ** 2 * np.identity(m) - K00 - sl - sr)) < 1e-9
    t = np.real(np.trace(g @ s.gamma(sl) @ g.conj().T @ s.gamma(sr)))
    assert abs(t - tm[2]) < 1e-8 * max(1.0, abs(t))


def test_sig_large_lead_block_uses_the_global_workspace():
    """lead blocks whose eleven m x m complex work matrices exceed shared memory (m > 35) run the same decimation on a per-CTA
    workspace in global memory: m = 60 against the oracle"""
    from sclmd_b200.selfenergy import sig
    m = 60
    K00, K11, K01 = P.chain_blocks(m, seed=9, k2=0.08)
    full = np.zeros((2 * m, 2 * m))
    full[:m, :m], full[m:, m:], full[:m, m:], full[m:, :m] = K00, K11, K01, K01.T
    s = sig(None, 0.06, range(0, m), range(m, 2 * m), dynmatfile=full, num=24, eta=1e-3)
    om = s.ep[[0, 5, 13, 24]]
    for d in ('L', 'R'):
        se = s.selfenergy_sweep(om, d)
        want = np.array([O.sig_selfenergy(K00, K11, s.K01, s.K10, w, s.eta, d) for w in om])
        assert relerr(se, want) < 1e-9
        assert np.array_equal(s.iterations, np.array([O.sig_sgf(K00, K11, s.K01, s.K10, w, s.eta, d)[1] for w in om]))
    tm = s.tm_sweep(om)
    wtm = np.array([O.sig_tm(K00, K11, s.K01, s.K10, w, s.eta) for w in om])
    assert np.max(np.abs(tm - wtm)) < 1e-8 * max(1.0, np.abs(wtm).max())
    assert relerr(s.sgf(float(om[2]), 'R'), O.sig_sgf(K00, K11, s.K01, s.K10, float(om[2]), s.eta, 'R')[0]) < 1e-9


def test_bpt_config3_full_sweep_properties():
    """size-independent properties at the full config-3 shape (n = 483, a few thousand frequencies, several batches and streams):
    reciprocity T_LR(w) = T_RL(w) (the two sweeps factorise differently ordered matrices), 0 <= T <= number of channels,
    batching invariance (any sub-range of the grid gives the same numbers), and agreement with the oracle on a sample"""
    from sclmd_b200.negf import bpt
    natoms = 201
    K = P.spring_chain_dyn(natoms, seed=14) / O.RPC ** 2
    fixed = [list(range(0, 60)), list(range(543, 603))]
    bath = [list(range(60, 210)), list(range(393, 543))]
    b = bpt(None, 0.25, 0.1, bath, fixed, dynmatfile=K, num=1000)
    om = np.linspace(0.0, 0.25 / O.RPC, 1500)
    t_lr = b.tm_sweep(om)
    b2 = bpt(None, 0.25, 0.1, [bath[1], bath[0]], fixed, dynmatfile=K, num=1000)
    t_rl = b2.tm_sweep(om)
    scale = max(1.0, np.abs(t_lr).max())
    assert np.max(np.abs(t_lr - t_rl)) < 1e-8 * scale
    assert t_lr.min() > -1e-10 and t_lr.max() <= 150.0 + 1e-9 and t_lr[0] == 0.0
    sub = b.tm_sweep(om[700:713])
    assert np.max(np.abs(sub - t_lr[700:713])) < 1e-10 * scale
    iL, iR = O.bpt_reduce_index(bath[0], 60), O.bpt_reduce_index(bath[1], 60)
    for k in (1, 333, 901, 1499):
        want = O.bpt_tm(b.dynmat, om[k], 0.1, iL, iR)
        assert abs(t_lr[k] - want) < 1e-8 * max(1.0, abs(want)), k


def test_bpt_singular_matrix_is_reported():
    """numpy.linalg.LinAlgError in the reference (negf.py:208): a structurally singular M(w) must come back as an error"""
    from sclmd_b200.negf import bpt
    from sclmd_b200._lib import SclmdError
    K = np.zeros((36, 36))
    K[17, 17] = -(1e-9 * 1e-9)                               # at w = 0, z^2 = -(1e-9)^2: row 17 of M = z^2 - K is exactly zero
    b = bpt(None, 0.25, 0.1, [list(range(3, 12)), list(range(24, 33))], [list(range(0, 3)), list(range(33, 36))], dynmatfile=K, num=4)
    with pytest.raises(SclmdError):
        b.tm_sweep(np.array([0.0]))


def test_bpt_green_functions_and_building_blocks():
    """bpt.retargf / advangf from the device (full inverse through the batched LU) and the self-energy building blocks, against
    numpy on the reference's formulas (negf.py:153-215), with and without the biased block"""
    from sclmd_b200.negf import bpt
    K = P.spring_chain_dyn(12, seed=80) / O.RPC ** 2
    fixed = [list(range(0, 3)), list(range(33, 36))]
    bath = [list(range(3, 12)), list(range(24, 33))]
    b = bpt(None, 0.25, 0.1, bath, fixed, dynmatfile=K, num=20)
    n = len(b.dynmat)

    def ref_gf(w, adv):
        sl, sr = b.retarselfenergy(w, bath[0]), b.retarselfenergy(w, bath[1])
        sb = b.retarbiasselfenergy(w, b.dofatomofbias)
        if adv:
            sl, sr = sl.conj().T, sr.conj().T
            sb = sb.conj().T if b.isbias else 0
        return np.linalg.inv((w + 1e-9j) ** 2 * np.identity(n) - b.dynmat - sl - sr - sb)
    for w in (3.0, 57.3, 211.0):
        assert relerr(b.retargf(w), ref_gf(w, False)) < 1e-9
        assert relerr(b.advangf(w), ref_gf(w, True)) < 1e-9
    g = b.gamma(b.retarselfenergy(57.3, bath[0]))
    assert np.allclose(np.diag(g)[:9], -2 * 57.3 / 0.1) and abs(g[12, 12]) == 0      # negf.py:214-215 with Sigma = -i w/damp
    t = np.real(np.trace(b.retargf(57.3) @ g @ b.retargf(57.3).conj().T @ b.gamma(b.retarselfenergy(57.3, bath[1]))))
    assert abs(t - b.tm(57.3)) < 1e-8 * max(1.0, abs(t))              # the reference's own tm formula (negf.py:240-242)
    bd, cp, cm = P.psd(6, 82, 2.0), P.sym(6, 83, 1.5), P.antisym(6, 84, 1.5)
    b.setbias(0.6, bdamp=bd, chiplus=cp, chiminus=cm, dofatomofbias=list(range(15, 21)))
    for w in (17.0, 140.0):
        gr, ga = b.retargf(w), b.advangf(w)
        assert relerr(gr, ref_gf(w, False)) < 1e-9 and relerr(ga, ref_gf(w, True)) < 1e-9
        want = w ** 2 * np.trace(np.real(np.linalg.multi_dot([gr, b.totalkselfenergy(w, 300.0), ga])[12:18][:, 12:18]))
        assert abs(b.ps(w, 300.0, list(range(15, 21))) - want) < 1e-8 * abs(want)     # negf.py:236 from the building blocks
    sweep = b.green_sweep(np.array([17.0, 140.0]))
    assert sweep.shape == (2, n, n) and relerr(sweep[1], ref_gf(140.0, False)) < 1e-9


@pytest.mark.parametrize("natoms,nfix,nb", [(4, 0, 1), (5, 3, 3), (7, 3, 2), (43, 3, 20)])
def test_bpt_small_and_odd_systems(natoms, nfix, nb):
    """smallest systems: n = 12 with one-dof leads, odd orders (the working matrix is padded by an identity dof), leads that are not
    whole blocks, n = 123 (two 64-column blocks, a partial one)"""
    from sclmd_b200.negf import bpt
    K = P.spring_chain_dyn(natoms, seed=3) / O.RPC ** 2
    n3 = 3 * natoms
    fixed = [list(range(0, nfix)), list(range(n3 - nfix, n3))]
    bath = [list(range(nfix, nfix + nb)), list(range(n3 - nfix - nb, n3 - nfix))]
    b = bpt(None, 0.25, 0.1, bath, fixed, dynmatfile=K, num=10)
    om = np.array([0.0, 1.3, 44.0, 171.0, 350.0])
    got = b.tm_sweep(om)
    iL, iR = O.bpt_reduce_index(bath[0], nfix), O.bpt_reduce_index(bath[1], nfix)
    want = np.array([O.bpt_tm(b.dynmat, w, 0.1, iL, iR) for w in om])
    assert np.max(np.abs(got - want)) < 1e-8 * max(1.0, np.abs(want).max())
    sel = list(range(nfix + nb, nfix + nb + 2))
    ps = b.ps_sweep(om[1:4], 300.0, sel)
    wantps = np.array([O.bpt_ps_nobias(b.dynmat, w, 300.0, 0.1, iL, iR, np.array(sel) - nfix) for w in om[1:4]])
    assert relerr(ps, wantps) < 1e-8


@pytest.mark.parametrize("natoms", [401, 700])
def test_bpt_systems_beyond_1024_dofs(natoms):
    """n = 1197 and n = 2094: the tall sub-panels are factorised 4 resp. 2 columns at a time (rows per thread x width is bounded by
    the register file); same pivoting, same result as the oracle's dense inverse"""
    from sclmd_b200.negf import bpt
    K = P.spring_chain_dyn(natoms, seed=16) / O.RPC ** 2
    n3 = 3 * natoms
    fixed = [list(range(0, 3)), list(range(n3 - 3, n3))]
    bath = [list(range(3, 153)), list(range(n3 - 153, n3 - 3))]
    b = bpt(None, 0.25, 0.1, bath, fixed, dynmatfile=K, num=10)
    om = np.array([2.0, 87.0, 260.0])
    got = b.tm_sweep(om)
    iL, iR = O.bpt_reduce_index(bath[0], 3), O.bpt_reduce_index(bath[1], 3)
    want = np.array([O.bpt_tm(b.dynmat, w, 0.1, iL, iR) for w in om])
    assert np.max(np.abs(got - want)) < 1e-8 * max(1.0, np.abs(want).max())


def test_bpt_handles_are_independent_and_own_their_state(tmp_path, monkeypatch):
    """the C-ABI handle of the sweeps (sclmd_bpt_create ... destroy): two junctions side by side, one with a bias block set and removed
    again, per-handle profiling; the matrix is uploaded once per handle and a changed matrix gives a new handle"""
    import ctypes as C
    from sclmd_b200 import _lib
    from sclmd_b200.negf import bpt
    monkeypatch.chdir(tmp_path)
    fixed = [list(range(0, 3)), list(range(33, 36))]
    bath = [list(range(3, 12)), list(range(24, 33))]
    Ka = P.spring_chain_dyn(12, seed=80) / O.RPC ** 2
    Kb = P.spring_chain_dyn(12, seed=91) / O.RPC ** 2
    a = bpt(None, 0.25, 0.1, bath, fixed, dynmatfile=Ka, num=20)
    b = bpt(None, 0.25, 0.2, bath, fixed, dynmatfile=Kb, num=20)
    om = np.linspace(5.0, 300.0, 17)
    iL, iR = O.bpt_reduce_index(bath[0], 3), O.bpt_reduce_index(bath[1], 3)
    wa = np.array([O.bpt_tm(a.dynmat, w, 0.1, iL, iR) for w in om])
    wb = np.array([O.bpt_tm(b.dynmat, w, 0.2, iL, iR) for w in om])
    L = _lib.lib()
    _lib.check(L.sclmd_bpt_set_profiling(a._handle(), 1))
    ta1, tb1, ta2 = a.tm_sweep(om), b.tm_sweep(om), a.tm_sweep(om)
    assert a._handle().value != b._handle().value
    assert relerr(ta1, wa) < 1e-8 and relerr(tb1, wb) < 1e-8 and np.array_equal(ta1, ta2)
    ms, n = (C.c_double * 7)(), (C.c_int64 * 7)()
    _lib.check(L.sclmd_bpt_get_profile(a._handle(), ms, n, None, None))
    assert n[0] == 2 and sum(ms) > 0                      # two sweeps of handle a, one k_build launch each
    _lib.check(L.sclmd_bpt_get_profile(b._handle(), ms, n, None, None))
    assert sum(n) == 0                                    # profiling is a property of the handle
    # bias block on / off on the same handle
    bd, cp, cm = P.psd(6, 82, 2.0), P.sym(6, 83, 1.5), P.antisym(6, 84, 1.5)
    h0 = a._handle().value
    a.setbias(0.6, bdamp=bd, chiplus=cp, chiminus=cm, dofatomofbias=list(range(15, 21)))
    tbias = a.tm_sweep(om)
    assert a._handle().value == h0 and relerr(tbias, wa) > 1e-3
    a.isbias = False
    assert np.array_equal(a.tm_sweep(om), ta1)
    with pytest.raises(_lib.SclmdError):                  # ps without Keldysh weights on a handle that carries a bias block
        _lib.check(L.sclmd_bpt_set_bias(a._handle(), 12, 6, _lib.dptr(bd), _lib.dptr(cp), _lib.dptr(cm), 0.3))
        nb = np.ones(len(om))
        sel = _lib.as_i32(range(12, 18))
        out = np.empty(len(om))
        _lib.check(L.sclmd_bpt_ps(a._h, _lib.dptr(om), _lib.dptr(nb), len(om), _lib.iptr(sel), len(sel), _lib.dptr(out)))
    # an edited matrix is seen (new handle), close() frees the device side and the next sweep comes back
    a.dynmat = a.dynmat * 1.01
    t3 = a.tm_sweep(om)
    assert relerr(t3, np.array([O.bpt_tm(a.dynmat, w, 0.1, iL, iR) for w in om])) < 1e-8
    a.close()
    assert a._h is None and np.array_equal(a.tm_sweep(om), t3)
    a.close()
    b.close()
