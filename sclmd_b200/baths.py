"""Electron and phonon baths with the reference's class names, constructor signatures and
attributes (sclmd/baths.py:55-458).  The arithmetic -- noise generation, memory-kernel
construction and the bath force inside the MD step -- runs on the device."""
import sys

import numpy as np

from . import _lib, noise as _noise
from ._lib import as_f64, check, dptr
from .functions import antisymmetrize, chkShape, flinterp_index, symmetrize


def exlist(a, indices):
    """baths.py:12-13"""
    return a[indices]


def _is_diagonal(m):
    m = np.asarray(m)
    return m.ndim == 2 and not np.any(m - np.diag(np.diagonal(m)))


def gamt(tl, wl, gwl, gam, eta_ad=0, device=0):
    """baths.py:19-52: gamma(w) -> gamma(t) by direct cosine transform on the wl grid.

    kernel[k] = 2*mean_i[ flinterp(wl_i, gwl, gam) cos(wl_i t_k) ] * wl[-1]/pi    (eta_ad == 0)."""
    print("eta=0" if eta_ad == 0 else "eta!=0")
    gam = as_f64(gam)
    wl = as_f64(wl)
    tl = as_f64(tl)
    shape = gam.shape[1:]
    m = int(np.prod(shape))
    g2 = gam.reshape(gam.shape[0], m)
    gi = np.empty((len(wl), m))
    for i, w in enumerate(wl):                       # interpolation indices are host-side (functions.py:117-143)
        i0, i1, wt = flinterp_index(w, gwl)
        gi[i] = g2[i0] if i0 == i1 else g2[i0] + wt * (g2[i0] - g2[i1])
    giT = np.ascontiguousarray(gi.T)
    out = np.empty((len(tl), m))
    if eta_ad == 0:
        check(_lib.lib().sclmd_gamt(device, len(tl), len(wl), m, dptr(tl), dptr(wl), dptr(giT), dptr(out)))
    else:   # baths.py:43-50: mean over the two damped exponentials = 2 * table; times wl[-1]/pi
        alpha = 2.0 / len(wl) * wl[-1] / np.pi
        check(_lib.lib().sclmd_cos_transform(device, len(tl), len(wl), m, dptr(tl), dptr(wl), dptr(giT), float(eta_ad), alpha, dptr(out)))
    return out.reshape((len(tl),) + shape)


class _BathBase:
    """What sclmd_b200.md.md needs from a bath: cids, ml, kernel, optional Mq/Mp, a noise plan."""
    _md = None
    _index = None
    _noise = None
    _noise_version = 0
    _device = None

    @property
    def device(self):
        """the GPU the bath's own device work (gmem, gnoi before AddBath) runs on: the md's device once attached, else this rank's"""
        if self._device is None:
            from . import parallel as PAR
            return PAR.local_device(0)
        return self._device

    @device.setter
    def device(self, value):
        self._device = None if value is None else int(value)

    @property
    def noise(self):
        return self._noise

    @noise.setter
    def noise(self, value):
        self._noise = None if value is None else np.asarray(value, dtype=float)
        self._noise_version += 1
        self._noise_seed = None           # an injected series: no Philox key regenerates it

    def _noise_row(self, t):
        nz = self.noise
        if nz is None:
            raise RuntimeError("bforce: the bath has no host copy of its noise (generated on the device); set bath.noise or call gnoi()")
        nz = np.asarray(nz)
        return nz[t % self.nmd] if nz.ndim == 2 else nz[0, t % self.nmd]

    def _engine_kernel(self):
        """(kernel, Mq, Mp) for MDEngine.add_bath: diagonal storage when every kernel[j] is diagonal."""
        k = np.asarray(self.kernel, dtype=float)
        if all(_is_diagonal(k[j]) for j in range(k.shape[0])):
            k = np.ascontiguousarray(np.diagonal(k, axis1=1, axis2=2))
        return k

    def _generate_device_noise(self, engine, index, traj0):
        from . import parallel as PAR
        # ONE Philox key for the whole ensemble (rank 0's draw); the counters carry the GLOBAL trajectory index, so a sharded
        # ensemble gets the same noise as the same ensemble on one GPU
        seed = int(np.random.randint(0, 2 ** 62))
        if self._md is not None and getattr(self._md, "sharded", False):
            seed = PAR.broadcast_int(seed)
        return self._regenerate_device_noise(engine, index, traj0, seed)

    def _regenerate_device_noise(self, engine, index, traj0, seed):
        """fill the device table from a known Philox key (also the resume path of md.Run: the key is stored in the checkpoint)"""
        plan = self._plan(engine.device)
        check(_lib.lib().sclmd_md_generate_noise(engine._h, index, plan._h, int(seed), int(traj0)))
        plan.close()
        self._noise_seed = int(seed)
        return seed


class ebath(_BathBase):
    """baths.py:55-255.  cats are DOF indices (baths.py:79-80)."""

    def __init__(self, cats, T, dt, nmd, wmax=None, nw=None, bias=0.,
                 efric=None, exim=None, exip=None, zeta1=None, zeta2=None, classical=False, zpmotion=True):
        self.cats = np.array(cats, dtype='int')
        self.cids = np.array(cats, dtype='int')
        self.nc = len(self.cids)
        self.T, self.wmax = T, wmax
        self.nw, self.bias = nw, bias
        self.dt, self.nmd = dt, nmd
        self.cur = np.zeros(nmd)
        self.classical = classical
        self.zpmotion = zpmotion
        self.wl = None if nw is None or wmax is None else [self.wmax * i / nw for i in range(nw)]
        self.CheckEmat(efric, exim, exip, zeta1, zeta2)
        self.ml = 1
        self.noise = None

    def CheckEmat(self, efric=None, exim=None, exip=None, zeta1=None, zeta2=None):
        """baths.py:100-174: symmetrise efric/exip/zeta1, antisymmetrise exim/zeta2."""
        if efric is None:
            print("ebath.CheckEmat: no efric provided, setting ebath to False")
            self.efric = self.kernel = self.exim = self.exip = self.zeta1 = self.zeta2 = None
            self.ebath = False
            return
        n = chkShape(efric)
        if n != self.nc:
            print("ebath.CheckEmat: efric shape error!")
            sys.exit()
        self.efric = symmetrize(efric)
        self.kernel = np.array([self.efric])
        self.exip = np.zeros(shape=(n, n))
        self.exim = np.zeros(shape=(n, n))
        self.zeta1 = np.zeros(shape=(n, n))
        self.zeta2 = np.zeros(shape=(n, n))
        self.ebath = True
        for name, val, fn in (("exim", exim, antisymmetrize), ("exip", exip, symmetrize), ("zeta1", zeta1, symmetrize),
                              ("zeta2", zeta2, antisymmetrize)):
            if val is not None:
                if chkShape(val) != self.nc:
                    print("ebath.CheckEmat: the dimension of %s is wrong!" % name)
                    sys.exit(0)
                setattr(self, name, fn(val))

    def _plan(self, device=0):
        return _noise.e_plan(self.efric, self.exim, self.exip, self.bias, self.T, self.wmax, self.dt, self.nmd,
                             self.classical, self.zpmotion, device)

    def gnoi(self):
        """baths.py:176-192.  Attached to an ensemble md (ntraj > 1) the table is generated on the
        device for every trajectory without a host copy; otherwise self.noise = one series [nmd, nc]."""
        if self.nmd is None:
            print("ebath.gnoi: nmd not set!")
            sys.exit()
        if self.dt is None:
            print("ebath.gnoi: dt not set!")
            sys.exit()
        if self.ebath is False:
            print("ebath.gnoi: ebath is False!")
            sys.exit()
        if self._md is not None and self._md._device_noise(self):
            return
        plan = self._plan(self.device)
        self.noise = plan.generate(1, int(np.random.randint(0, 2 ** 62)))[0]
        plan.close()

    def bforce(self, t, phis, qhis):
        """baths.py:224-255 as a stand-alone call (the time loop itself never calls it: the force is fused into the device step).
        The matrix products run on the device (sclmd_dgemm_nt)."""
        from .noise import mf
        pc, qc = np.asarray(phis)[0][self.cids], np.asarray(qhis)[0][self.cids]
        f = self._noise_row(t) - _lib.dgemm_nt(pc[None], np.asarray(self.efric, dtype=float), 1.0, self.device)[0]
        ex = [np.asarray(m, dtype=float) if m is not None else None for m in (self.exim, self.zeta1, self.zeta2)]
        if all(m is not None and np.any(m) for m in ex):            # baths.py:233: only if all three have a non-zero element
            f = f + self.bias * _lib.dgemm_nt(qc[None], ex[0] - ex[1], 1.0, self.device)[0]
            f = f - self.bias * _lib.dgemm_nt(pc[None], ex[2], 1.0, self.device)[0]
        return mf(f, self.cids, len(np.asarray(phis)[0]))

    def _engine_extra(self):
        """baths.py:233: the exim / zeta1 / zeta2 forces act only if ALL THREE have a non-zero entry."""
        if self.exim.any() and self.zeta1.any() and self.zeta2.any():
            return self.bias * (self.exim - self.zeta1), -self.bias * self.zeta2
        return None, None

    def GetSig(self):
        """baths.py:194-209 (host set-up helper)."""
        if self.wl is None:
            print("ebath.GetSig:wl is not set")
            sys.exit()
        nc = chkShape(self.efric)
        self.sig = np.zeros((len(self.wl), nc, nc), complex)
        for i, w in enumerate(self.wl):
            self.sig[i] = -1.j * w * (self.efric + self.bias * self.zeta2) + self.bias * self.zeta1 - self.bias * self.exim

    def SetMDsteps(self, dt, nmd):
        self.dt, self.nmd = dt, nmd

    def setbias(self, bias=0.0):
        self.bias = bias
        print("ebath.setbias: WARNING--BIAS CHANGED! YOU NEED TO REGENERATE THE NOISE!")


class phbath(_BathBase):
    """baths.py:258-458."""

    def __init__(self, T, cats, debye, nw, dt, nmd, ml=None, mcof=2.0, sig=None, gamma=None, gwl=None, K00=None,
                 K01=None, V01=None, eta_ad=0, classical=False, zpmotion=True):
        self.classical = classical
        self.zpmotion = zpmotion
        self.T, self.debye, self.cats = T, debye, np.array(cats, dtype='int')
        self.K00, self.K01, self.V01 = K00, K01, V01
        self.dt, self.nmd, self.ml = dt, nmd, ml
        self.kernel = None
        self.cids = np.array(cats, dtype='int')
        self.nc = len(self.cids)
        self.wmax = mcof * debye
        self.local = False
        self.nw = nw
        self.wl = [self.wmax * i / nw for i in range(nw)]
        self.gamma = gamma
        self.sig = sig
        self.gwl = gwl
        self.cur = np.zeros(nmd)
        self.eta_ad = eta_ad
        self.noise = None
        if self.UseK():
            print("phbath: Calculating self-energy is not implemented yet.")
            sys.exit(0)
        elif self.UseG() or self.UsePi():
            if self.UsePi():
                if len(self.sig[0]) != self.nc:
                    print("phbath: inconsist between cids and sig!")
                    sys.exit()
                self.ggamma()
            if self.UseG():
                if len(self.gamma[0]) != self.nc:
                    print("phbath: inconsist between cids and gamma!")
                    sys.exit()
        else:
            # Debye model, Adelman & Doll JCP 64, 2375 (1976)  (baths.py:333-340)
            phfric = debye * np.pi / 6.0
            self.gamma = np.array([np.diag(phfric + np.zeros(int(self.nc)))])
            self.gwl = np.array([0])
            self.local = True
            self.ml = 1

    def SetMDsteps(self, dt, nmd):
        self.dt, self.nmd = dt, nmd

    def SetMemlen(self, len):
        self.ml = len

    def SetT(self, T):
        self.T = T

    def UseG(self):
        return self.gamma is not None and self.gwl is not None

    def UsePi(self):
        return self.sig is not None and self.gwl is not None

    def UseK(self):
        return self.K00 is not None and self.K01 is not None and self.V01 is not None

    def ggamma(self):
        """baths.py:375-395: gamma(w) = -Im Sigma(w)/w, the w == 0 node copied from the next one."""
        if self.sig is None:
            print("phbath.Gamma: self.sig is not set, need it to calculate gamma")
            sys.exit()
        a = []
        for i in range(len(self.gwl)):
            if self.gwl[i] == 0:
                a.append(-np.imag(self.sig[i + 1]) / self.gwl[i + 1])
            else:
                a.append(-np.imag(self.sig[i]) / self.gwl[i])
        self.gamma = np.array(a)

    def _plan(self, device=0):
        return _noise.ph_plan(self.gamma, self.gwl, self.T, self.wmax, self.dt, self.nmd, self.classical, self.zpmotion, device)

    def gnoi(self):
        """baths.py:397-409."""
        if self.dt is None or self.nmd is None:
            print("phbath.gnoi: the md information dt and nmd are not set!")
            sys.exit()
        if self._md is not None and self._md._device_noise(self):
            return
        plan = self._plan(self.device)
        self.noise = plan.generate(1, int(np.random.randint(0, 2 ** 62)))[0]
        plan.close()

    def gmem(self):
        """baths.py:412-446: memory kernel in the time domain."""
        if self.ml is None or self.dt is None:
            print("phbath.gmem: length of memory kernel not set!")
            sys.exit()
        if self.local:
            self.ml = 1
            self.kernel = self.gamma
        else:
            tl = [self.dt * i for i in range(self.ml)]
            self.kernel = np.real(gamt(tl, self.wl, self.gwl, self.gamma, self.eta_ad, self.device))
            if self.eta_ad != 0:
                # baths.py:429-445: gamma is re-derived from the damped kernel, gamma(w_i) = dt sum_t kernel(t) cos(w_i t)
                k2 = np.ascontiguousarray(self.kernel.reshape(self.ml, -1).T)
                gw = as_f64(self.gwl)
                tt = as_f64(tl)
                out = np.empty((len(gw), k2.shape[0]))
                check(_lib.lib().sclmd_cos_transform(self.device, len(gw), len(tt), k2.shape[0], dptr(gw), dptr(tt), dptr(k2), 0.0,
                                                     float(self.dt), dptr(out)))
                self.gammaOld = self.gamma
                self.gamma = out.reshape((len(gw),) + self.kernel.shape[1:])

    def bforce(self, t, phis, qhis):
        """baths.py:448-458 as a stand-alone call: noise - sum_i kernel[i] . phis[i][cids] (* dt when ml > 1), contracted on the device"""
        from .noise import mf
        kern = np.asarray(self.kernel, dtype=float)
        ml = kern.shape[0]
        hist = np.asarray(phis)[:ml][:, self.cids].reshape(1, -1)                       # [1, ml*nc]
        fric = _lib.dgemm_nt(hist, np.ascontiguousarray(kern.transpose(1, 0, 2)).reshape(self.nc, -1), 1.0, self.device)[0]
        f = self._noise_row(t) - (fric * self.dt if ml > 1 else fric)
        return mf(f, self.cids, len(np.asarray(phis)[0]))

    def _engine_extra(self):
        return None, None
