"""Ballistic phonon transport by NEGF with the reference's class `bpt` (sclmd/negf.py:8-277).
The frequency sweeps (gettm, getps) run on the device; LAMMPS is optional: pass the dynamical
matrix as `dynmatfile=<path or ndarray>` together with `natoms=`."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import as_f64, as_i32, check, dptr, iptr


class bpt:
    def __init__(self, infile, maxomega, damp, dofatomofbath, dofatomfixed=[[], []], dynmatfile=None, num=1000,
                 natoms=None, device=None, write_files=None):
        self.rpc = 6.582119569e-4     # reduced Planck constant, eV*ps
        # falsefrequencies.dat / omegas.dat / eigvecs.dat in the working directory (negf.py:100-102): by default whenever the
        # dynamical matrix comes from a file, as in the reference's flow; not when the caller hands over an array
        self.write_files = (not isinstance(dynmatfile, np.ndarray)) if write_files is None else bool(write_files)
        self.bc = 8.617333262e-5      # Boltzmann constant, eV/K
        self.damp = damp              # ps
        self.maxomega = maxomega / self.rpc
        self.intnum = num
        self.dofatomfixed = [list(dofatomfixed[0]), list(dofatomfixed[1])]
        self.isbias = False
        self.dofatomofbias = []
        self.dofatomofbath = [list(dofatomofbath[0]), list(dofatomofbath[1])]
        self.dynmatfile = dynmatfile
        if device is None:                 # the GPU of this rank under torchrun
            from . import parallel as _PAR
            device = _PAR.local_device(0)
        self.device = device
        self.natoms = natoms
        self._h, self._hkey = None, None
        self.getdynmat(infile)

    def setbias(self, bias, bdamp=None, chiplus=None, chiminus=None, dofatomofbias=[]):
        """negf.py:27-37: attach a biased electron bath on the contiguous block dofatomofbias[0]..dofatomofbias[-1]"""
        np.seterr(divide='ignore', invalid='ignore')
        self.isbias = True
        self.bias = bias / self.rpc
        self.biasgamma = bdamp
        self.chiplus = chiplus
        self.chiminus = chiminus
        self.dofatomofbias = list(dofatomofbias)
        if len(self.biasgamma) != len(self.chiminus) or len(self.biasgamma) != len(self.chiplus) or \
                len(self.biasgamma) != len(self.dofatomofbias):
            raise ValueError('Bias parameters not set correctly')

    def _bias_block(self):
        t1, t2 = self.dofatomofbias[0], self.dofatomofbias[-1] + 1          # negf.py:165-166
        b0 = t1 - len(self.dofatomfixed[0])
        nb = t2 - t1
        mats = [as_f64(np.ascontiguousarray(m), (nb, nb)) for m in (self.biasgamma, self.chiplus, self.chiminus)]
        return b0, nb, mats

    def _handle(self):
        """the device handle of this junction (sclmd_bpt_create): K uploaded once, workspace and streams kept between sweeps;
        re-created when the matrix, the leads or the damping change, the bias block re-set before every sweep (O(nb^2) host copy)"""
        iL, iR = self._reduced(self.dofatomofbath[0]), self._reduced(self.dofatomofbath[1])
        k = self.dynmat
        key = (k.shape, float(k.sum()), float(np.abs(k).sum()), float(self.damp), iL.tobytes(), iR.tobytes(), self.device)
        if self._h is None or key != self._hkey:
            self.close()
            h = C.c_void_p()
            check(_lib.lib().sclmd_bpt_create(self.device, len(k), dptr(as_f64(k)), iptr(iL), len(iL), iptr(iR), len(iR), float(self.damp),
                                              C.byref(h)))
            self._h, self._hkey = h, key
        if self.isbias:
            b0, nb, (bd, cp, cm) = self._bias_block()
            check(_lib.lib().sclmd_bpt_set_bias(self._h, b0, nb, dptr(bd), dptr(cp), dptr(cm), float(self.bias)))
        else:
            check(_lib.lib().sclmd_bpt_set_bias(self._h, 0, 0, None, None, None, 0.0))
        return self._h

    def close(self):
        """frees the device side (sclmd_bpt_destroy); the next sweep creates it again"""
        if getattr(self, "_h", None) is not None:
            _lib.lib().sclmd_bpt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _keldysh_weights(self, om, T):
        """per-frequency scalars of totalkselfenergy (negf.py:177-193), Bose factors evaluated with bpt.bosedist"""
        kd, kr1, kr2, ki = (np.zeros(len(om)) for _ in range(4))
        b = self.bias
        with np.errstate(all="ignore"):
            for i, w in enumerate(om):
                n0 = float(self.bosedist(w, T))
                cp = (w + b) * (2 * float(self.bosedist(w + b, T)) - 2 * n0)
                cm = (w - b) * (2 * float(self.bosedist(w - b, T)) - 2 * n0)
                kd[i] = 2.0 * w / self.damp * n0
                kr1[i] = w * 2 * n0
                kr2[i] = (cp + cm) / 2
                ki[i] = -b * 2 * n0 + (cm - cp) / 2
        return kd, kr1, kr2, ki

    def getdynmat(self, infile):
        """negf.py:39-102 without the LAMMPS dependency: the dynamical matrix comes from
        `dynmatfile` (text file as written by LAMMPS `dynamical_matrix`, or an ndarray)."""
        if self.dynmatfile is None:
            raise RuntimeError("bpt: computing the dynamical matrix needs LAMMPS (negf.py:40-63), which is outside this "
                               "build; pass dynmatfile=<file or ndarray>")
        if isinstance(self.dynmatfile, np.ndarray):
            dyn = np.array(self.dynmatfile, dtype=float)
        else:
            dat = np.loadtxt(self.dynmatfile)
            dynlen = int(3 * np.sqrt(len(dat) / 3))
            dyn = dat.reshape((dynlen, dynlen))
        if self.natoms is None:
            self.natoms = dyn.shape[0] // 3
        if dyn.shape[0] != self.natoms * 3:
            raise ValueError('System DOF test failed after load dynmat, check again')
        dyn = (dyn + dyn.transpose()) / 2
        fixed = list(self.dofatomfixed[0]) + list(self.dofatomfixed[1])
        dyn = np.delete(np.delete(dyn, fixed, axis=0), fixed, axis=1)
        self.dynmat = np.ascontiguousarray(dyn)
        if self.natoms * 3 != len(self.dofatomfixed[0]) + len(self.dofatomfixed[1]) + len(self.dynmat):
            raise ValueError('System DOF test failed, check again')
        eigvals, self.eigvecs = np.linalg.eigh(self.dynmat)
        self.omegas = [np.sqrt(v) * self.rpc if v > 0 else -np.sqrt(-v) * self.rpc for v in eigvals]
        ffi = [i for i, v in enumerate(eigvals) if not v > 0]
        print('%i false frequencies exist in %i frequencies' % (len(ffi), len(self.omegas)))
        if self.write_files:                              # negf.py:100-102
            np.savetxt('falsefrequencies.dat', ffi, fmt='%d')
            np.savetxt('omegas.dat', self.omegas)
            np.savetxt('eigvecs.dat', self.eigvecs)

    def _reduced(self, dofs):
        """negf.py:195-204: bath dofs are in the unreduced 3N numbering; the leading fixed block is removed"""
        idx = np.asarray(list(dofs), dtype=np.int64) - len(self.dofatomfixed[0])
        if len(idx) and (idx.min() < 0 or idx.max() >= len(self.dynmat)):
            raise ValueError('System DOF test failed, check again')
        return as_i32(idx)

    def bosedist(self, omega, T):
        """negf.py:217-226"""
        if abs(T) < 1e-30:
            return 1 / (np.exp(self.rpc * omega * np.iinfo(np.int32).max) - 1)
        elif abs(omega / T) < 1e-30:
            return np.iinfo(np.int32).max
        return 1 / (np.exp(self.rpc * omega / self.bc / T) - 1)

    # ---- the reference's building blocks (negf.py:153-215).  The sweeps above never form these matrices; they are kept for
    #      scripts that call them directly.  Self-energies are O(n^2) assembly on the host, Green functions come from the device.
    def cleanse(self, semat):
        """negf.py:195-204"""
        semat = np.delete(semat, self.dofatomfixed[0], axis=0)
        semat = np.delete(semat, self.dofatomfixed[0], axis=1)
        semat = np.delete(semat, [dof - len(self.dofatomfixed[0]) for dof in self.dofatomfixed[1]], axis=0)
        semat = np.delete(semat, [dof - len(self.dofatomfixed[0]) for dof in self.dofatomfixed[1]], axis=1)
        if len(semat) != len(self.dynmat) or self.natoms * 3 != len(self.dofatomfixed[0]) + len(self.dofatomfixed[1]) + len(semat):
            raise ValueError('System DOF test failed, check again')
        return semat

    def retarselfenergy(self, omega, dofatoms):
        """negf.py:153-157"""
        semat = np.zeros((self.natoms * 3, self.natoms * 3), dtype=np.complex128)
        for dofatom in dofatoms:
            semat[dofatom, dofatom] = -1j * omega / self.damp
        return self.cleanse(semat)

    def advanselfenergy(self, omega, dofatoms):
        return self.retarselfenergy(omega, dofatoms).conjugate().transpose()

    def retarbiasselfenergy(self, omega, dofatoms):
        """negf.py:162-172"""
        if self.isbias:
            semat = np.zeros((self.natoms * 3, self.natoms * 3), dtype=np.complex128)
            t1, t2 = dofatoms[0], dofatoms[-1] + 1
            semat[t1:t2, t1:t2] = -1j * omega * np.asarray(self.biasgamma) - self.bias * np.asarray(self.chiminus)
            return self.cleanse(semat)
        return 0

    def advanbiasselfenergy(self, omega, dofatoms):
        r = self.retarbiasselfenergy(omega, dofatoms)
        return r.conjugate().transpose() if self.isbias else 0

    def kselfenergy(self, omega, T, dofatoms):
        """negf.py:177-178"""
        return -2 * np.imag(self.retarselfenergy(omega, dofatoms)) * self.bosedist(omega, T)

    def kbiasselfenergy(self, omega, T, dofatoms):
        """negf.py:180-190"""
        if self.isbias:
            semat = np.zeros((self.natoms * 3, self.natoms * 3), dtype=np.complex128)
            t1, t2 = dofatoms[0], dofatoms[-1] + 1
            cp, cm = np.asarray(self.chiplus), np.asarray(self.chiminus)
            with np.errstate(all="ignore"):
                semat[t1:t2, t1:t2] = ((cp - 1j * cm) * (omega + self.bias) * (2 * self.bosedist(omega + self.bias, T) - 2 * self.bosedist(omega, T))
                                       + (cp + 1j * cm) * (omega - self.bias) * (2 * self.bosedist(omega - self.bias, T) - 2 * self.bosedist(omega, T))) / 2
            return (1j * self.retarbiasselfenergy(omega, dofatoms)) * 2 * self.bosedist(omega, T) + self.cleanse(semat)
        return 0

    def totalkselfenergy(self, omega, T):
        """negf.py:192-193"""
        return (self.kselfenergy(omega, T, self.dofatomofbath[0]) + self.kselfenergy(omega, T, self.dofatomofbath[1])
                + self.kbiasselfenergy(omega, T, self.dofatomofbias))

    def gamma(self, Pi):
        """negf.py:214-215"""
        return -1j * (Pi - Pi.conjugate().transpose())

    def green_sweep(self, omegas, advanced=False):
        """full Green functions G(w) [len(omegas), n, n] from the device (batched LU with the identity as right-hand side)"""
        om = as_f64(np.atleast_1d(omegas))
        n = len(self.dynmat)
        out = np.empty((len(om), n, n), dtype=np.complex128)
        check(_lib.lib().sclmd_bpt_green(self._handle(), dptr(om), len(om), 1 if advanced else 0, out.ctypes.data_as(_lib.c_double_p)))
        return out

    def retargf(self, omega):
        """negf.py:206-208"""
        return self.green_sweep([omega])[0]

    def advangf(self, omega):
        """negf.py:210-212 (the +1e-9j of z is kept, as the reference does)"""
        return self.green_sweep([omega], advanced=True)[0]

    def tm_sweep(self, omegas):
        """T(w) for an array of frequencies (ps^-1) on the device (negf.py:240-242 for each)"""
        om = as_f64(omegas)
        out = np.empty(len(om))
        check(_lib.lib().sclmd_bpt_tm(self._handle(), dptr(om), len(om), dptr(out)))
        return out

    def tm(self, omega):
        return float(self.tm_sweep(np.array([omega], dtype=float))[0])

    def gettm(self, vector=False):
        """negf.py:104-119"""
        from . import parallel as PAR
        x = np.linspace(0, self.maxomega, self.intnum + 1)
        # inside a torch.distributed job the grid is cut into contiguous blocks over the ranks and all-gathered (frequencies are independent)
        self.tmnumber = np.array(np.column_stack((x, PAR.sharded_sweep(self.tm_sweep, x))))
        if PAR.rank_world()[0] == 0:
            np.savetxt('transmission.dat', np.column_stack((self.tmnumber[:, 0] * self.rpc, self.tmnumber[:, 1])))

    def ps_sweep(self, omegas, T, atomlist):
        om = as_f64(omegas)
        sel = self._reduced(atomlist)
        out = np.empty(len(om))
        h = self._handle()
        if self.isbias:   # negf.py:234-236
            kd, kr1, kr2, ki = self._keldysh_weights(om, T)
            check(_lib.lib().sclmd_bpt_ps_bias(h, dptr(om), dptr(kd), dptr(kr1), dptr(kr2), dptr(ki), len(om), iptr(sel), len(sel), dptr(out)))
            return out
        with np.errstate(all="ignore"):
            nb = np.array([float(self.bosedist(w, T)) for w in om])
        check(_lib.lib().sclmd_bpt_ps(h, dptr(om), dptr(nb), len(om), iptr(sel), len(sel), dptr(out)))
        return out

    def ps(self, omega, T, atomlist):
        """negf.py:228-238"""
        return float(self.ps_sweep(np.array([omega], dtype=float), T, atomlist)[0])

    def getps(self, T, maxomega, intnum, atomlist=None, filename=None, vector=False, omegalist=None):
        """negf.py:121-150"""
        if atomlist is None:
            atomlist = np.array(range(0, len(self.dynmat))) + len(self.dofatomfixed[0])
        x2 = np.sort(omegalist) / self.rpc if omegalist is not None else np.linspace(0, maxomega / self.rpc, intnum + 1)
        from . import parallel as PAR
        self.psnumber = np.array(np.column_stack((x2, PAR.sharded_sweep(lambda om: self.ps_sweep(om, T, atomlist), x2))))
        name = 'powerspectrum.' + (str(filename) + '.' if filename is not None else '') + str(T) + '.dat'
        if PAR.rank_world()[0] == 0:
            np.savetxt(name, np.column_stack((self.psnumber[:, 0] * self.rpc, self.psnumber[:, 1])))

    def thermalcurrent(self, T, delta):
        """negf.py:245-270: trapezoid over the stored transmission, nW"""
        n = len(self.tmnumber[:, 0]) - 1
        if n != self.intnum:
            raise ValueError('Error in number of omega')
        arr = np.array([self.rpc * w / 2 / np.pi * t * (self.bosedist(w, T * (1 + 0.5 * delta)) - self.bosedist(w, T * (1 - 0.5 * delta)))
                        for w, t in self.tmnumber])
        return (float(self.tmnumber[-1, 0] - self.tmnumber[0, 0]) / n / 2.) * (2 * arr.sum() - arr[0] - arr[-1]) * 1.60217662 * 1e2

    def thermalconductance(self, T, delta):
        return self.thermalcurrent(T, delta) / (T * delta)

    def thermalconductivity(self, T, delta, L, A):
        return self.thermalconductance(T, delta) * L / A * 10
