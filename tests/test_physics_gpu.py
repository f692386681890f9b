"""End-to-end physics cross-check of the two pipelines (the pairing of examples/runmd.py and examples/runnegf.py): the thermal
conductance from the ensemble-averaged heat current of quantum-thermostat MD equals the Landauer value from the NEGF transmission
of the same harmonic junction, within the statistical error of the MD average.  Exercises the device noise generator (quantum
spectra), the ensemble time stepping, md.Run's bookkeeping and the batched LU sweep in one go."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_md_conductance_matches_landauer(tmp_path, monkeypatch):
    from sclmd_b200.md import md
    from sclmd_b200.baths import ebath
    from sclmd_b200.negf import bpt
    from sclmd_b200 import units as U
    monkeypatch.chdir(tmp_path)
    np.random.seed(12345)
    ntraj, nmd, T, delta, dt = 1024, 4096, 300.0, 0.5, 0.25 / 0.658
    natoms = 101
    lap = 2.0 * np.eye(natoms) - np.eye(natoms, k=1) - np.eye(natoms, k=-1)
    K = np.kron(lap, np.diag([0.010, 0.006, 0.003])) + 1e-6 * np.eye(3 * natoms)      # ordered chain, three polarisations
    fixed = [list(range(0, 30)), list(range(273, 303))]
    cats = [list(range(30, 105)), list(range(198, 273))]
    damp = 100 / 0.658211814201041                                                    # 0.1 ps (examples/runmd.py:45, runnegf.py:17)
    m = md(dt, nmd, T, axyz=[["C", float(i), 0.0, 0.0] for i in range(natoms)], dyn=K, nstart=0, nstop=2, ntraj=ntraj)
    for b, Tb in enumerate((T * (1 + delta / 2), T * (1 - delta / 2))):
        m.AddBath(ebath(cats[b], Tb, dt, nmd, wmax=1.0, nw=500, efric=np.identity(75) / damp))
    m.AddConstr(fixed)
    m.Run()                                   # run 0 equilibrates, run 1 is measured
    j = np.array([np.asarray(b.cur).reshape(ntraj, nmd).mean(axis=1) * U.curcof for b in m.baths])     # nW per trajectory
    jm = (j[0] - j[1]) / 2
    kappa_md, err = jm.mean() / (T * delta), jm.std(ddof=1) / np.sqrt(ntraj) / (T * delta)
    RPC = 6.582119569e-4
    nb = bpt(None, 0.25, 0.1, cats, fixed, dynmatfile=np.array(m.dyn) / RPC ** 2, num=2000)
    nb.gettm()
    kappa_negf = nb.thermalconductance(T, delta)
    assert 0.3 < kappa_negf < 1.5                                    # three channels of at most one conductance quantum (0.28 nW/K at 300 K)
    assert err < 0.1 * kappa_negf                                    # the average is resolved
    assert abs(kappa_md - kappa_negf) < 4.0 * err, (kappa_md, err, kappa_negf)
    assert abs((j[0] + j[1]).mean()) < 0.1 * abs(jm.mean())          # stationary: what enters on the left leaves on the right
