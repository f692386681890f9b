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
    """functions.py:221-236: sum_dof |Fourier1D(p)|^2 / (dt nmd)"""
    pst = np.transpose(np.array(ps))
    if nmd != pst.shape[1]:
        raise ValueError("power: ps shape error!")
    dw = 2. * np.pi / dt / nmd
    psw = np.fft.ifft(pst, axis=1) * (2. * np.pi / dw)
    psw = np.real(np.transpose(psw * np.conjugate(psw)))
    return np.array([[i * dw, np.sum(psw[i]) / dt / nmd] for i in range(nmd)])
