"""Host-side scalar helpers with the reference's names and edge-case behaviour
(sclmd/functions.py).  They only prepare spectral weights / interpolation indices for the
device kernels; no hot-path arithmetic runs here."""
import sys

import numpy as np

from . import units as U

np.seterr(over="ignore")


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
