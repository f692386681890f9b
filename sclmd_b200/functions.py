"""Host-side scalar helpers with the reference's names and edge-case behaviour
(sclmd/functions.py).  They only prepare spectral weights / interpolation indices for the
device kernels; no hot-path arithmetic runs here."""
import sys

import numpy as np

from . import units as U

np.seterr(over="ignore")


def coth(x):
    """functions.py:59-67"""
    if x == 0.0:
        print("coth:coth(0) is infinity")
        sys.exit(0)
    return np.cosh(x) / np.sinh(x)


def xcoth(x):
    """functions.py:70-77"""
    return 1.0 if x == 0.0 else x * np.cosh(x) / np.sinh(x)


def device_fft(a, sign, scale, device=0):
    """out[..., m] = scale * sum_k a[..., k] exp(sign 2 pi i k m / n) along the last axis, on the device (sclmd_fft)"""
    from . import _lib
    a = np.asarray(a)
    n = a.shape[-1]
    re = np.ascontiguousarray(a.real, dtype=np.float64).reshape(-1, n)
    im = np.ascontiguousarray(a.imag, dtype=np.float64).reshape(-1, n) if np.iscomplexobj(a) else None
    ore, oim = np.empty_like(re), np.empty_like(re)
    _lib.check(_lib.lib().sclmd_fft(int(device), int(n), int(re.shape[0]), _lib.dptr(re), _lib.dptr(im), int(sign), float(scale),
                                    _lib.dptr(ore), _lib.dptr(oim)))
    return (ore + 1j * oim).reshape(a.shape)


class myfft:
    """functions.py:11-53 with the transforms on the device (lengths of the form 2^a 3^b 5^c)."""

    def __init__(self, dt, n, device=0):
        self.dt = dt
        self.N = n
        self.dw = 2 * np.pi / dt / n
        self.device = device

    def Fourier1D(self, a):
        """f(j) = dt sum_i f(i) e^(+I 2pi i j/N)  ==  numpy.fft.ifft(a) * 2 pi / dw"""
        if len(a) != self.N:
            print("MyFFT.Fourier1D: array length error!")
            sys.exit(0)
        return device_fft(np.asarray(a), +1, (2. * np.pi / self.dw) / self.N, self.device)

    def iFourier1D(self, a):
        """f(i) = dw/2pi sum_j f(j) e^(-I 2pi i j/N)  ==  numpy.fft.fft(a) * dw / 2 pi"""
        if len(a) != self.N:
            print("MyFFT.iFourier1D: array length error!")
            sys.exit(0)
        return device_fft(np.asarray(a), -1, self.dw / 2 / np.pi, self.device)


def bose(w, T):
    """functions.py:80-99 (bose(0,T>0) = 0 by fiat; T == 0 branches)."""
    if T == 0.0:
        if w == 0.0:
            return 1 / (np.exp(1.0 / U.kb) - 1)
        elif w < 0.0:
            return -1.0
        return 0.0
    if w == 0.0:
        return 0.0
    return 1.0 / (np.exp(w / U.kb / T) - 1.0)


def fermi(ep, mu, T):
    """functions.py:102-114."""
    if T == 0.0:
        if ep < mu:
            return 1.0
        elif ep > mu:
            return 0.0
        return 0.5
    return 1 / (np.exp((ep - mu) / U.kb / T) + 1)


def nearest(b, bs):
    """functions.py:137-143: index of the FIRST element of bs nearest to b."""
    return int(np.argmin(np.abs(np.array(bs) - b)))


def flinterp_index(x, xs):
    """Index/weight form of flinterp (functions.py:117-134):
    value = ys[i0] + w*(ys[i0]-ys[i1]); flat end half-intervals give (i,i,0)."""
    i = nearest(x, xs)
    if i == len(xs) - 1 or i == 0:
        return i, i, 0.0
    dd = x - xs[i]
    if dd < 0:
        return i, i - 1, dd / (xs[i] - xs[i - 1])
    return i, i + 1, dd / (xs[i] - xs[i + 1])


def flinterp(x, xs, ys):
    """functions.py:117-134."""
    i0, i1, w = flinterp_index(x, xs)
    if i0 == i1:
        return ys[i0]
    return ys[i0] + w * (ys[i0] - ys[i1])


def chkShape(a):
    """functions.py:166-176."""
    ash = np.shape(np.array(a))
    if ash[0] == ash[1]:
        return ash[0]
    print("The matrix should be a n by n matrix")
    sys.exit(0)


def symmetrize(a):
    aa = np.array(a)
    return 0.5 * (aa + np.transpose(aa))


def antisymmetrize(a):
    aa = np.array(a)
    return 0.5 * (aa - np.transpose(aa))


def dagger(a):
    aa = np.array(a)
    if aa.shape[0] != aa.shape[1]:
        print("Not sqaure matrix")
        sys.exit(0)
    return np.transpose(np.conjugate(aa))


def hermitianize(a):
    aa = np.array(a)
    return 0.5 * (aa + dagger(aa))


def powerspecq(qs, dt, nmd, device=0):
    """functions.py:203-218: (dw i)^2 sum_dof |Fourier1D(q)|^2 / (dt nmd); the batched transform runs on the device"""
    qst = np.transpose(np.array(qs))
    if nmd != qst.shape[1]:
        print("power: qs shape error!")
        sys.exit()
    dw = 2. * np.pi / dt / nmd
    qsw = device_fft(qst, +1, (2. * np.pi / dw) / nmd, device)
    tot = np.sum(np.real(qsw * np.conjugate(qsw)), axis=0)
    return np.array([[i * dw, (dw * i) ** 2 * tot[i] / dt / nmd] for i in range(nmd)])


def powerspecp(ps, dt, nmd, device=0):
    """functions.py:221-236: sum_dof |Fourier1D(p)|^2 / (dt nmd); the batched transform runs on the device"""
    pst = np.transpose(np.array(ps))
    if nmd != pst.shape[1]:
        print("power: ps shape error!")
        sys.exit()
    dw = 2. * np.pi / dt / nmd
    psw = device_fft(pst, +1, (2. * np.pi / dw) / nmd, device)
    tot = np.sum(np.real(psw * np.conjugate(psw)), axis=0)
    return np.array([[i * dw, tot[i] / dt / nmd] for i in range(nmd)])


def mm(*args):
    """functions.py:159-164: chained matrix product, every factor on the device (sclmd_dgemm_nt; real matrices)"""
    from . import _lib
    tmp = np.array(args[0], dtype=float)
    for mat in args[1:]:
        m = np.asarray(mat, dtype=float)
        vec_l, vec_r = tmp.ndim == 1, m.ndim == 1
        r = _lib.dgemm_nt(tmp.reshape(1, -1) if vec_l else tmp, m.reshape(1, -1) if vec_r else np.ascontiguousarray(m.T))
        tmp = r.reshape(-1) if vec_l or vec_r else r
        if vec_l and vec_r:
            tmp = tmp[0]
    return tmp


def rpadleft(bs, b):
    """functions.py:146-153 (kept for API parity; the device keeps ring buffers instead)."""
    if len(bs) > 1:
        return np.concatenate((np.array([b]), np.array(bs)[:-1]), axis=0)
    elif len(bs) == 1:
        return np.array([b])
    print("len(bs) is less than 1")
    sys.exit()


def mdot(*args):
    return np.linalg.multi_dot([im for im in args])
