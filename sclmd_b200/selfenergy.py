"""Surface self-energy of semi-infinite leads with the reference's class `sig`
(sclmd/selfenergy.py:7-178).  Sancho-Rubio decimation and the transmission run on the device."""
import numpy as np

from . import _lib
from ._lib import as_f64, check, dptr


class sig:
    def __init__(self, infile, maxomega, atomgroup0, atomgroup1, dofatomfixed=[[], []], dynmatfile=None, num=1000,
                 eta=0.164e-3, device=None):
        self.rpc = 6.582119569e-4
        self.maxomega = maxomega / self.rpc
        self.intnum = num
        self.eta = eta / self.rpc
        self.dofatomK00 = list(atomgroup0)
        self.dofatomK11 = list(atomgroup1)
        self.dofatomfixed = dofatomfixed
        self.dynmatfile = dynmatfile
        if device is None:                 # the GPU of this rank under torchrun
            from . import parallel as _PAR
            device = _PAR.local_device(0)
        self.device = device
        self.ep = np.linspace(0, self.maxomega, self.intnum + 1)
        self.getdynmat(infile)
        self.getdk()

    def getdynmat(self, infile):
        """selfenergy.py:28-91 without LAMMPS: dynmatfile = text file or ndarray (full 3N x 3N)"""
        if self.dynmatfile is None:
            raise RuntimeError("sig: computing the dynamical matrix needs LAMMPS (selfenergy.py:29-53), which is outside "
                               "this build; pass dynmatfile=<file or ndarray>")
        if isinstance(self.dynmatfile, np.ndarray):
            self.dynmat = np.array(self.dynmatfile, dtype=float)
        else:
            dat = np.loadtxt(self.dynmatfile)
            dynlen = int(3 * np.sqrt(len(dat) / 3))
            self.dynmat = dat.reshape((dynlen, dynlen))
        self.natoms = self.dynmat.shape[0] // 3
        # frequencies of the system without its fixed dofs (selfenergy.py:66-91); the three text files are written whenever the
        # matrix comes from a file, as in the reference's flow
        fixed = list(self.dofatomfixed[0]) + list(self.dofatomfixed[1])
        red = np.delete(np.delete(self.dynmat, fixed, axis=0), fixed, axis=1)
        eigvals, eigvecs = np.linalg.eigh(red)
        self.omegas = [np.sqrt(v) * self.rpc if v > 0 else -np.sqrt(-v) * self.rpc for v in eigvals]
        ffi = [i for i, v in enumerate(eigvals) if not v > 0]
        print('%i false frequencies exist in %i frequencies' % (len(ffi), len(self.omegas)))
        if not isinstance(self.dynmatfile, np.ndarray):
            np.savetxt('falsefrequencies.dat', ffi, fmt='%d')
            np.savetxt('omegas.dat', self.omegas)
            np.savetxt('eigvecs.dat', eigvecs)

    def getdk(self):
        """selfenergy.py:93-103"""
        self.K00 = self.dynmat[self.dofatomK00, :][:, self.dofatomK00]
        self.K11 = self.dynmat[self.dofatomK11, :][:, self.dofatomK11]
        self.K01 = self.dynmat[self.dofatomK00, :][:, self.dofatomK11]
        self.K10 = self.dynmat[self.dofatomK11, :][:, self.dofatomK00]
        err = np.amax(abs(self.K01 - np.transpose(self.K10))) / np.amax(abs(self.K01))
        if err > 1e-8:
            raise ValueError('Error: K01 and K10 are not symmetric', err)
        self.K01 = (self.K01 + np.transpose(self.K10)) / 2
        self.K10 = np.transpose(self.K01)

    def _k(self):
        return [as_f64(np.ascontiguousarray(m)) for m in (self.K00, self.K11, self.K01, self.K10)]

    def selfenergy_sweep(self, omegas, direction):
        if direction not in ('R', 'L'):
            raise ValueError('Wrong direction, should only be R or L')
        om = as_f64(omegas)
        m = len(self.K00)
        k00, k11, k01, k10 = self._k()
        out = np.empty((len(om), m, m), dtype=np.complex128)
        iters = np.zeros(len(om), dtype=np.int32)
        check(_lib.lib().sclmd_sig_selfenergy(self.device, m, dptr(k00), dptr(k11), dptr(k01), dptr(k10), float(self.eta),
                                              direction.encode(), dptr(om), len(om), out.ctypes.data_as(_lib.c_double_p),
                                              iters.ctypes.data_as(_lib.c_int32_p)))
        self.iterations = iters
        return out

    def sgf_sweep(self, omegas, direction):
        """surface Green function of one lead for every frequency (selfenergy.py:105-131 per frequency)"""
        if direction not in ('R', 'L'):
            raise ValueError('Wrong direction, should only be R or L')
        om = as_f64(np.atleast_1d(omegas))
        m = len(self.K00)
        k00, k11, k01, k10 = self._k()
        out = np.empty((len(om), m, m), dtype=np.complex128)
        iters = np.zeros(len(om), dtype=np.int32)
        check(_lib.lib().sclmd_sig_sgf(self.device, m, dptr(k00), dptr(k11), dptr(k01), dptr(k10), float(self.eta),
                                       direction.encode(), dptr(om), len(om), out.ctypes.data_as(_lib.c_double_p),
                                       iters.ctypes.data_as(_lib.c_int32_p)))
        self.iterations = iters
        return out

    def sgf(self, omega, direction):
        """selfenergy.py:105-131"""
        return self.sgf_sweep([omega], direction)[0]

    def selfenergy(self, omega, direction):
        """selfenergy.py:133-140"""
        return self.selfenergy_sweep(np.array([omega], dtype=float), direction)[0]

    def gamma(self, Pi):
        return -1j * (Pi - Pi.conjugate().transpose())

    def tm_sweep(self, omegas):
        om = as_f64(omegas)
        m = len(self.K00)
        k00, k11, k01, k10 = self._k()
        out = np.empty(len(om))
        check(_lib.lib().sclmd_sig_tm(self.device, m, dptr(k00), dptr(k11), dptr(k01), dptr(k10), float(self.eta), dptr(om),
                                      len(om), dptr(out)))
        return out

    def green_sweep(self, omegas):
        """G(w) = inv((w + 1e-8 i)^2 - K00 - Sigma_L - Sigma_R) for every frequency, on the device"""
        om = as_f64(np.atleast_1d(omegas))
        m = len(self.K00)
        k00, k11, k01, k10 = self._k()
        out = np.empty((len(om), m, m), dtype=np.complex128)
        check(_lib.lib().sclmd_sig_green(self.device, m, dptr(k00), dptr(k11), dptr(k01), dptr(k10), float(self.eta), dptr(om),
                                         len(om), out.ctypes.data_as(_lib.c_double_p)))
        return out

    def retargf(self, omega):
        """selfenergy.py:145-147"""
        return self.green_sweep([omega])[0]

    def tm(self, omega):
        """selfenergy.py:149-151"""
        return float(self.tm_sweep(np.array([omega], dtype=float))[0])

    def getse(self, direction):
        """selfenergy.py:153-166"""
        from . import parallel as PAR
        se = PAR.sharded_sweep(lambda om: self.selfenergy_sweep(om, direction), self.ep)      # frequency blocks over the ranks, all-gathered
        dosx = -np.trace(np.imag(se), axis1=1, axis2=2) * self.ep / np.pi
        self.dos = np.array(np.column_stack((self.ep, dosx)))
        if PAR.rank_world()[0] == 0:
            np.savetxt('densityofstates_' + str(direction) + '.dat', np.column_stack((self.dos[:, 0] * self.rpc, self.dos[:, 1])))
        return se

    def gettm(self):
        """selfenergy.py:168-178"""
        from . import parallel as PAR
        self.tmnumber = np.array(np.column_stack((self.ep, PAR.sharded_sweep(self.tm_sweep, self.ep))))
        if PAR.rank_world()[0] == 0:
            np.savetxt('transmission.dat', np.column_stack((self.tmnumber[:, 0] * self.rpc, self.tmnumber[:, 1])))
