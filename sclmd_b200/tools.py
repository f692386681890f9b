"""Post-processing with the reference's function names (sclmd/tools.py:132-215,
sclmd/functions.py:203-236): run-averaged heat flux and thermal conductance from the
kappa.* files written by md.Run(), power spectra from saved velocities.  Host-side file
processing -- the step AFTER the hot path ("next" rows of the scope table)."""
import glob

import numpy as np


def _read_kappa(bathnum):
    temperture = None
    for filename in glob.glob('kappa.*.bath0.run0.dat'):
        with open(filename, 'r') as f:
            for line in f:
                temperture = float(line.split()[1])
    times = int(len(glob.glob('kappa.*.bath0.run*.dat')))
    kb = np.empty([bathnum, times])
    for i in range(bathnum):
        for j in range(times):
            for files in glob.glob("kappa." + str(int(temperture)) + ".bath" + str(i) + ".run" + str(j) + ".dat"):
                with open(files, 'r') as f:
                    for line in f:
                        kb[i][j] = line.split()[2]
    return temperture, kb


def calHF(dlist=1, bathnum=2):
    """tools.py:132-163: cumulative average heat flux per bath, first `dlist` runs dropped"""
    temperture, kb = _read_kappa(bathnum)
    oldkb = np.delete(kb, list(range(dlist)), axis=1)
    balancekb = np.array(oldkb)
    for i in range(balancekb.shape[0]):
        for j in range(balancekb.shape[1]):
            balancekb[i][j] = np.mean(oldkb[i][0:j + 1])
    np.savetxt('heatflux.' + str(int(temperture)) + '.dat', np.transpose(balancekb))


def calTC(delta, dlist=1, bathnum=2, L=None, A=None):
    """tools.py:166-215: kappa = (J0 - J1)/2/(delta*T), mean and std over runs"""
    temperture, kb = _read_kappa(bathnum)
    dl = list(range(dlist))
    if delta != 0:
        if bathnum == 2:
            kappa = np.delete((kb[0] - kb[1]) / 2 / (delta * temperture), dl)
        elif bathnum == 3:
            kappa = np.delete((kb[0] + kb[1] - kb[2]) / 4 / (delta * temperture), dl)
        np.savetxt('thermalconductance.' + str(int(temperture)) + '.dat', (np.mean(kappa), np.std(kappa)),
                   header="Mean(nW/K) Std(nW/K)")
        if L is not None and A is not None:
            np.savetxt('thermalconductivity.' + str(int(temperture)) + '.dat',
                       (np.mean(kappa * L / A * 10), np.std(kappa * L / A * 10)), header="Mean(W/m-K) Std(W/m-K)")
    if bathnum == 2:
        kappa = np.delete((kb[0] - kb[1]) / 2, dl)
    elif bathnum == 3:
        kappa = np.delete(-(kb[0] + kb[1] - kb[2]) / 4, dl)
    np.savetxt('heatflux-between-baths.' + str(int(temperture)) + '.dat', (np.mean(kappa), np.std(kappa)), header="Mean(nW) Std(nW)")


def powerspecp(ps, dt, nmd):
    """functions.py:221-236 (kept importable from tools as md.GetPower expects): device transform, see functions.powerspecp"""
    from .functions import powerspecp as _p
    return _p(ps, dt, nmd)


def phbath_from_sig(s, direction, T, cats, nw, dt, nmd, ml, mcof=2.0, debye=None, eta_ad=0, classical=False, zpmotion=True):
    """The step BEFORE the hot path: a `sig` lead self-energy sweep (selfenergy.py:153-166, frequencies in ps^-1, Sigma in
    ps^-2) becomes the `phbath(sig=..., gwl=...)` of an MD run (baths.py:294,322-327,375-395), whose energies are in eV:
    hbar w -> rpc * w and Sigma -> rpc^2 * Sigma with rpc = 6.582119569e-4 eV ps.  `debye` defaults to the end of the
    self-energy grid divided by `mcof`, so that the kernel transform (`gmem` -> `gamt`) and the noise cutoff
    (wmax = mcof * debye, baths.py:305-308,408) cover exactly the tabulated range.  Call `.gmem()` on the result before
    stepping, as with the reference (md.Run never does, baths.py:412)."""
    from .baths import phbath
    se = np.asarray(s.getse(direction))                 # device sweep; also writes densityofstates_<direction>.dat
    gwl = np.asarray(s.ep, dtype=float) * s.rpc
    sig_md = se * s.rpc ** 2
    if len(cats) != se.shape[1]:
        raise ValueError("phbath_from_sig: %d bath dofs but the self-energy is %dx%d" % (len(cats), se.shape[1], se.shape[2]))
    if debye is None:
        debye = float(gwl[-1]) / mcof
    return phbath(T, cats, debye, nw, dt, nmd, ml=ml, mcof=mcof, sig=sig_md, gwl=gwl, eta_ad=eta_ad, classical=classical,
                  zpmotion=zpmotion)


def get_atomname(mass):
    """tools.py:218-226: element whose tabulated mass lies within 0.01 amu of `mass` (None if there is none)"""
    from . import units as U
    for name, m in U.AtomicMassTable.items():
        if abs(mass - m) < 0.01:
            return name
    return None


def get_atommass(name):
    """tools.py:229-237: tabulated mass of an element (None for an unknown name)"""
    from . import units as U
    return U.AtomicMassTable.get(name)


def eff(dynmatfilename='dynmat.dat'):
    """tools.py:240-259, 'eliminate false frequencies': symmetrise the dynamical matrix of the text file, set negative eigenvalues
    to zero and rebuild, until none is left; writes 'mod<file>' and returns the matrix"""
    dat = np.loadtxt(dynmatfilename)
    n = int(3 * np.sqrt(len(dat) / 3))
    dyn = dat.reshape((n, n))
    dyn = (dyn + dyn.T) / 2
    w, v = np.linalg.eigh(dyn)
    while not (w > 0).all():
        for i in np.nonzero(w < 0)[0]:
            print('False frequency exists in system DOF %i ' % i)
        w = np.where(w < 0, 0.0, w)
        dyn = (v * w) @ np.linalg.inv(v)
        dyn = (dyn + dyn.T) / 2
        w, v = np.linalg.eigh(dyn)
    np.savetxt('mod' + dynmatfilename, dyn)
    return dyn


def avdf(dffiles=["deltaforce.run0.npy"], outputname="deltaforce", abs=False):
    """tools.py:7-32: running mean and standard deviation of the force differences md.CompareForce recorded
    (deltaforce.run<j>.npy), over the first 1, 2, ... files; writes <outputname>-mean<i>.dat and -deviation<i>.dat"""
    chunks = [np.load(f) for f in dffiles]
    per = len(chunks[0])
    allv = np.concatenate(chunks, axis=0)
    if abs:
        allv = np.abs(allv)
    for i in range(len(dffiles)):
        part = allv[0:int((i + 1) * per)]
        mean = np.mean(part, axis=0)
        np.savetxt(outputname + "-mean" + str(i) + ".dat", mean)
        np.savetxt(outputname + "-deviation" + str(i) + ".dat", np.sqrt(np.mean((part - mean) ** 2, axis=0)))
