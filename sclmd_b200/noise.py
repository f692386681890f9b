"""Quantum coloured-noise generation on the device with the reference's function names
(sclmd/noise.py).  `phnoise` / `enoise` keep the reference signatures and return ONE series
[nmd, nc] by default; `ntraj`, `seed`, `xi`, `device` are optional extras."""
import ctypes as C

import numpy as np

from . import _lib, units as U
from ._lib import as_f64, check, dptr, iptr
from .functions import bose, chkShape, flinterp_index


def equ(w, cut, T, classical=False, zpmotion=True):
    """noise.py:249-270: 2*hw*(zp + n_B(hw)) below the cutoff, 2 kB T if classical or hw == 0."""
    hw = U.hbar * w
    zp = 0.5 if zpmotion is True else 0.0
    if hw < cut:
        if classical:
            return 2.0 * U.kb * T
        if hw == 0:
            return 2.0 * U.kb * T
        return 2.0 * hw * (zp + bose(hw, T))
    return 0.0


def nonequm(w, bias, T, classical=False):
    """noise.py:209-226"""
    hw1, hw2, small = U.hbar * w - bias, U.hbar * w, 10e-20
    if classical:
        hw1 = small if hw1 == 0. else hw1
        hw2 = small if hw2 == 0. else hw2
        return 2.0 * hw1 * (U.kb * T / hw1 - U.kb * T / hw2)
    return 2.0 * hw1 * (bose(hw1, T) - bose(hw2, T))


def nonequp(w, bias, T, classical=False):
    """noise.py:229-246"""
    hw1, hw2, small = U.hbar * w + bias, U.hbar * w, 10e-20
    if classical:
        hw1 = small if hw1 == 0. else hw1
        hw2 = small if hw2 == 0. else hw2
        return 2.0 * hw1 * (U.kb * T / hw1 - U.kb * T / hw2)
    return 2.0 * hw1 * (bose(hw1, T) - bose(hw2, T))


def phnoisew(gamma, wl, T, phcut, classical=False, zpmotion=True):
    """noise.py:28-46: the phonon noise spectrum equ(w) gamma(w) on the given grid (assembly only, no random numbers)"""
    gamma = np.array(gamma)
    return np.array([equ(wl[i], phcut, T, classical, zpmotion) * gamma[i] for i in range(len(wl))])


def enoisew(wl, efric, exim, exip, bias, T, ecut, classical=False, zpmotion=True):
    """noise.py:103-146: the Hermitian electron noise spectrum on the given grid (assembly only, no random numbers)"""
    from .functions import hermitianize
    efric, exim, exip = np.array(efric), np.array(exim), np.array(exip)
    out = np.zeros((len(wl),) + efric.shape, dtype=complex)
    for i, w in enumerate(wl):
        aw = equ(w, ecut, T, classical, zpmotion)
        awm = equ(U.hbar * w - bias, ecut, T, classical, zpmotion)
        awp = equ(U.hbar * w + bias, ecut, T, classical, zpmotion)
        amat = aw * efric + (-0.5 * aw * exip + 0.5 * awm * (exip + 1j * exim)) + (-0.5 * aw * exip + 0.5 * awp * (exip - 1j * exim))
        out[i] = hermitianize(amat)
    return out


def mf(f, cats, lens):
    """noise.py:15-22 scatter (host helper kept for API parity)."""
    t = np.zeros(lens)
    t[np.asarray(cats, dtype=int)] = f
    return t


class NoisePlan:
    """Per-frequency covariance factors held on the device (sclmd_noise_plan_* in the C ABI)."""

    def __init__(self, nmd, dt, nc, basis, idx, cre, cim=None, device=0):
        basis = as_f64(basis)
        nbasis = basis.shape[0]
        if basis.shape[1:] != (nc, nc):
            raise ValueError("basis must be [nbasis,%d,%d]" % (nc, nc))
        nw = nmd // 2 + 1
        idx = np.ascontiguousarray(idx, dtype=np.int32).reshape(nw, -1)
        nterm = idx.shape[1]
        cre = as_f64(cre, (nw, nterm))
        cim = None if cim is None else as_f64(cim, (nw, nterm))
        self.nmd, self.dt, self.nc, self.nw, self.device = int(nmd), float(dt), int(nc), nw, device
        self._h = C.c_void_p()
        check(_lib.lib().sclmd_noise_plan_create(device, self.nmd, self.dt, self.nc, nbasis, dptr(basis), nterm, iptr(idx),
                                                 dptr(cre), dptr(cim), C.byref(self._h)))

    def close(self):
        if self._h:
            _lib.lib().sclmd_noise_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def is_complex(self):
        return bool(check(_lib.lib().sclmd_noise_plan_is_complex(self._h)))

    def factors(self):
        n = self.nw * self.nc * self.nc
        if self.is_complex:
            out = np.empty(2 * n)
            check(_lib.lib().sclmd_noise_plan_get_factors(self._h, dptr(out)))
            return out.view(np.complex128).reshape(self.nw, self.nc, self.nc)
        out = np.empty(n)
        check(_lib.lib().sclmd_noise_plan_get_factors(self._h, dptr(out)))
        return out.reshape(self.nw, self.nc, self.nc)

    def set_factors(self, L):
        L = np.asarray(L)
        if L.shape != (self.nw, self.nc, self.nc):
            raise ValueError("factors must be [%d,%d,%d]" % (self.nw, self.nc, self.nc))
        if np.iscomplexobj(L):
            buf = np.ascontiguousarray(L, dtype=np.complex128).view(np.float64)
            check(_lib.lib().sclmd_noise_plan_set_factors(self._h, dptr(buf), 1))
        else:
            buf = as_f64(L)
            check(_lib.lib().sclmd_noise_plan_set_factors(self._h, dptr(buf), 0))

    def generate(self, ntraj=1, seed=0, traj0=0, xi=None):
        """[ntraj, nmd, nc] real series.  xi: injected standard normals [ntraj, nmd/2+1, nc]."""
        out = np.empty((ntraj, self.nmd, self.nc))
        if xi is not None:
            xi = as_f64(xi, (ntraj, self.nw, self.nc))
        check(_lib.lib().sclmd_noise_plan_generate(self._h, int(ntraj), dptr(xi), C.c_uint64(int(seed) & (2 ** 64 - 1)),
                                                   int(traj0), dptr(out)))
        return out

    def launch_count(self):
        return int(_lib.lib().sclmd_noise_plan_launch_count(self._h))

    def profile(self):
        """device ms of the last generate call per stage, and of the factorisation at plan creation"""
        ms = np.zeros(6)
        check(_lib.lib().sclmd_noise_plan_get_profile(self._h, dptr(ms)))
        return dict(draws_ms=float(ms[0]), gemm_ms=float(ms[1]), transform_ms=float(ms[2]), factor_ms=float(ms[3]),
                    n_cholesky=int(ms[4]), n_jacobi=int(ms[5]))


def ph_plan(gamma, wl, T, phcut, dt, nmd, classical=False, zpmotion=True, device=0):
    """Spectral weights of noise.py:73-79 as a device plan: A(w_i) = (dt nmd) equ(w_i) flinterp(w_i, wl, gamma)."""
    if nmd % 2:
        raise _lib.SclmdError("MyFFT.iFourier1D: array length error!")
    gamma = as_f64(gamma)
    nc = gamma.shape[1]
    nw = nmd // 2 + 1
    dw = 2.0 * np.pi / dt / nmd
    delta = dt * nmd
    idx = np.full((nw, 2), -1, dtype=np.int32)
    cre = np.zeros((nw, 2))
    for i in range(nw):
        w = dw * i
        s = delta * equ(w, phcut, T, classical, zpmotion)
        i0, i1, wt = flinterp_index(w, wl)
        idx[i, 0], cre[i, 0] = i0, s * (1.0 + wt) if i0 != i1 else s
        if i0 != i1:
            idx[i, 1], cre[i, 1] = i1, -s * wt
    return NoisePlan(nmd, dt, nc, gamma, idx, cre, None, device)


def e_plan(efric, exim, exip, bias, T, ecut, dt, nmd, classical=False, zpmotion=True, device=0):
    """noise.py:171-186: A = aw*efric + (-aw + awm/2 + awp/2)*exip + i*(awm-awp)/2*exim."""
    if nmd % 2:
        raise _lib.SclmdError("MyFFT.iFourier1D: array length error!")
    efric, exim, exip = as_f64(efric), as_f64(exim), as_f64(exip)
    nc = chkShape(efric)
    nw = nmd // 2 + 1
    dw = 2.0 * np.pi / dt / nmd
    delta = dt * nmd
    idx = np.tile(np.array([0, 1, 2], dtype=np.int32), (nw, 1))
    cre, cim = np.zeros((nw, 3)), np.zeros((nw, 3))
    for i in range(nw):
        w = dw * i
        aw = delta * equ(w, ecut, T, classical, zpmotion)
        awm = delta * equ(U.hbar * w - bias, ecut, T, classical, zpmotion)
        awp = delta * equ(U.hbar * w + bias, ecut, T, classical, zpmotion)
        cre[i] = (aw, -aw + 0.5 * awm + 0.5 * awp, 0.0)
        cim[i] = (0.0, 0.0, 0.5 * (awm - awp))
    if not exim.any():
        cim = None
    if not exip.any():
        idx[:, 1] = -1
    if cim is None:
        idx[:, 2] = -1
    return NoisePlan(nmd, dt, nc, np.stack([efric, exip, exim]), idx, cre, cim, device)


def _seed(seed):
    return int(np.random.randint(0, 2 ** 62)) if seed is None else int(seed)


def phnoise(gamma, wl, T, phcut, dt, nmd, classical=False, zpmotion=True, ntraj=None, seed=None, xi=None, device=0):
    """noise.py:50-100.  Returns [nmd, nc] (or [ntraj, nmd, nc] when ntraj is given)."""
    plan = ph_plan(gamma, wl, T, phcut, dt, nmd, classical, zpmotion, device)
    out = plan.generate(1 if ntraj is None else ntraj, _seed(seed), 0, xi)
    plan.close()
    return out[0] if ntraj is None else out


def enoise(efric, exim, exip, bias, T, ecut, dt, nmd, classical=False, zpmotion=True, ntraj=None, seed=None, xi=None, device=0):
    """noise.py:149-206.  Returns the real part the reference keeps (baths.py:191)."""
    plan = e_plan(efric, exim, exip, bias, T, ecut, dt, nmd, classical, zpmotion, device)
    out = plan.generate(1 if ntraj is None else ntraj, _seed(seed), 0, xi)
    plan.close()
    return out[0] if ntraj is None else out


def vargau(eval, evec, cof=1.0, device=0):
    """noise.py:273-305: one multivariate Gaussian draw from an eigen-decomposed covariance, V . r with r_k ~ N(0, sqrt(cof lambda_k))
    for the positive eigenvalues and 0 otherwise.  The normal deviates come from np.random.normal in the reference's order (drawn
    only for positive eigenvalues), so a seeded script reproduces the reference's numbers; the product runs on the device.
    (The bath classes do not call this: their noise is generated by the device generator, sclmd_noise_plan_*.)"""
    import sys
    nvec = np.array(evec)
    if nvec.ndim != 2 or len(eval) != nvec.shape[0] or nvec.shape[0] != nvec.shape[1]:
        print("vargau: shape error")
        sys.exit(0)
    rval = np.array([np.random.normal(0.0, np.sqrt(cof * v)) if cof * v > 0 else 0.0 for v in eval])
    if np.iscomplexobj(nvec):
        return _lib.dgemm_nt(rval[None, :], nvec.real, 1.0, device)[0] + 1j * _lib.dgemm_nt(rval[None, :], nvec.imag, 1.0, device)[0]
    return _lib.dgemm_nt(rval[None, :], nvec, 1.0, device)[0]
