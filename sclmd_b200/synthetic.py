"""Seeded synthetic junctions of SURVEY.md section 8d ("Synthetic inputs"): mass-weighted spring networks, the projection
md.setDyn applies to them, and model memory kernels.  Pure functions of their seeds; used by bench.py, by the parity tests
(tests/problems.py re-exports them) and by anyone who wants to try the package without LAMMPS."""
import numpy as np


def spring_chain_dyn(natoms, seed, kmax=0.04, onsite=1e-6):
    """Mass-weighted spring network on a quasi-1D ribbon: every atom bonds to
    the next two atoms with random stiffness; returns symmetric PSD K [3n,3n]
    with max eigenvalue ~ kmax (hbar*omega_max ~ 0.2 eV for kmax=0.04)."""
    rng = np.random.default_rng(seed)
    n = 3 * natoms
    K = np.zeros((n, n))
    pos = np.cumsum(rng.uniform(0.8, 1.2, size=(natoms, 3)), axis=0)
    for i in range(natoms):
        for j in (i + 1, i + 2):
            if j >= natoms:
                continue
            e = pos[j] - pos[i]
            e /= np.linalg.norm(e)
            e = e + 0.35 * rng.standard_normal(3)      # give the bond some transverse stiffness
            kb = rng.uniform(0.5, 1.0)
            blk = kb * np.outer(e, e)
            si, sj = slice(3 * i, 3 * i + 3), slice(3 * j, 3 * j + 3)
            K[si, si] += blk
            K[sj, sj] += blk
            K[si, sj] -= blk
            K[sj, si] -= blk
    K = 0.5 * (K + K.T) + onsite * np.eye(n)
    K *= kmax / np.linalg.eigvalsh(K).max()
    return K


def psd_project(K):
    """What md.setDyn does to the matrix it is given (md.py:264-281)."""
    K = 0.5 * (K + K.T)
    av, au = np.linalg.eigh(K)
    av = np.where(av < 0, 0.0, av)
    return au @ np.diag(av) @ au.T


def psd_project_modes(K):
    """psd_project together with the decomposition md.setDyn keeps (md.py:266-281): (K, lam clipped at 0, U)"""
    K = 0.5 * (K + K.T)
    av, au = np.linalg.eigh(K)
    av = np.where(av < 0, 0.0, av)
    return au @ np.diag(av) @ au.T, av, au


def full_kernel(ml, nc, dt, seed, gamma0=0.02, tau=12.0, w0=0.15, eps=0.1):
    """kernel[j] = gamma0 exp(-j dt/tau) cos(w0 j dt) (I + eps S), S symmetric seeded."""
    rng = np.random.default_rng(seed)
    S = rng.standard_normal((nc, nc))
    S = 0.5 * (S + S.T) / np.sqrt(nc)
    j = np.arange(ml)
    s = gamma0 * np.exp(-j * dt / tau) * np.cos(w0 * j * dt)
    return s[:, None, None] * (np.eye(nc) + eps * S)[None]


def diag_kernel(ml, nc, dt, seed, gamma0=0.02, tau=12.0, w0=0.15):
    rng = np.random.default_rng(seed)
    amp = gamma0 * rng.uniform(0.5, 1.5, size=nc)
    j = np.arange(ml)
    return (np.exp(-j * dt / tau) * np.cos(w0 * j * dt))[:, None] * amp[None, :]
