"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain NumPy restatement of the sclmd generalized-Langevin / NEGF hot path
(SURVEY.md section 8a).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and
only as the checker / the timed CPU baseline -- never on the product path.

Pinning: the reference ships no tests or golden vectors,
so this oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: the fixtures
in ``tests/golden/*.npz`` were written by ``oracle/make_golden.py`` which
imports ``/root/reference/sclmd`` in place and runs ``md.vv``, ``phnoise``,
``enoise``, ``gamt``, ``bpt.tm`` and ``sig.*`` on seeded inputs.
``tests/test_oracle_golden.py`` replays them through this file.

Every function cites the reference file:line it restates.
"""
import math

import numpy as np

# sclmd/units.py:5-10
HBAR = 1.0
KB = 0.000086173423
CURCOF = 243414.0
# sclmd/negf.py:13-15, selfenergy.py:12
RPC = 6.582119569e-4
BC = 8.617333262e-5


# --------------------------------------------------------------------------
# scalar helpers
# --------------------------------------------------------------------------
def bose(w, T):
    """functions.py:80-99."""
    if T == 0.0:
        if w == 0.0:
            return 1 / (np.exp(1.0 / KB) - 1)
        elif w < 0.0:
            return -1.0
        return 0.0
    if w == 0.0:
        return 0.0
    with np.errstate(over="ignore"):
        return 1.0 / (np.exp(w / KB / T) - 1.0)


def equ(w, cut, T, classical=False, zpmotion=True):
    """noise.py:249-270 -- spectral weight 2*hw*(zp + n_B)."""
    hw = HBAR * w
    zp = 0.5 if zpmotion is True else 0.0
    if hw < cut:
        if classical:
            return 2.0 * KB * T
        if hw == 0:
            return 2.0 * KB * T
        return 2.0 * hw * (zp + bose(hw, T))
    return 0.0


def nearest(b, bs):
    """functions.py:137-143 -- FIRST index of the minimum |bs-b|."""
    bst = np.abs(np.asarray(bs, dtype=float) - b)
    return int(np.argmin(bst))  # argmin returns the first minimal index, like list.index(min)


def flinterp(x, xs, ys):
    """functions.py:117-134 -- nearest-node linear interpolation, flat end half-intervals."""
    i = nearest(x, xs)
    if i == len(xs) - 1:
        return ys[-1]
    if i == 0:
        return ys[0]
    dd = x - xs[i]
    if dd < 0:
        return ys[i] + dd / (xs[i] - xs[i - 1]) * (ys[i] - ys[i - 1])
    return ys[i] + dd / (xs[i] - xs[i + 1]) * (ys[i] - ys[i + 1])


def flinterp_index(x, xs):
    """Index/weight form of flinterp: returns (i0, i1, w) with value = ys[i0] + w*(ys[i0]-ys[i1]).

    This is what the CUDA kernels consume; indices must match the reference bit-exactly."""
    i = nearest(x, xs)
    if i == len(xs) - 1 or i == 0:
        return i, i, 0.0
    dd = x - xs[i]
    if dd < 0:
        return i, i - 1, dd / (xs[i] - xs[i - 1])
    return i, i + 1, dd / (xs[i] - xs[i + 1])


def hermitianize(a):
    """functions.py:198-200."""
    a = np.asarray(a)
    return 0.5 * (a + a.conj().T)


# --------------------------------------------------------------------------
# memory kernel  gamma(w) -> gamma(t)
# --------------------------------------------------------------------------
def gamt(tl, wl, gwl, gam, eta_ad=0):
    """baths.py:19-52.  kernel[k] = 2*mean_i[flinterp(wl_i,gwl,gam) cos(wl_i t_k)]*wl[-1]/pi."""
    wl = np.asarray(wl, dtype=float)
    gi = np.array([np.asarray(flinterp(w, gwl, gam)) for w in wl])  # [nw, nc, nc]
    out = []
    if eta_ad == 0:
        for t in tl:
            c = np.cos(wl * t)
            out.append(2.0 * np.tensordot(c, gi, axes=(0, 0)) / len(wl) * wl[-1] / np.pi)
    else:
        for t in tl:
            with np.errstate(divide="ignore", invalid="ignore"):
                f = (wl / (wl - 1j * eta_ad) * np.exp(-1j * wl * t - eta_ad * t)
                     + wl / (wl + 1j * eta_ad) * np.exp(+1j * wl * t - eta_ad * t))
            out.append(np.tensordot(f, gi, axes=(0, 0)) / len(wl) * wl[-1] / np.pi)
    return np.array(np.real(out))


def ggamma(sig, gwl):
    """baths.py:375-395 -- gamma(w) = -Im Sigma(w)/w, w==0 entry copied from the next node."""
    a = []
    for i in range(len(gwl)):
        if gwl[i] == 0:
            a.append(-np.imag(sig[i + 1]) / gwl[i + 1])
        else:
            a.append(-np.imag(sig[i]) / gwl[i])
    return np.array(a)


# --------------------------------------------------------------------------
# coloured noise
# --------------------------------------------------------------------------
def vargau(ev, evec, draw):
    """noise.py:273-305.  ``draw(scale)`` stands for np.random.normal(0, scale):
    it is called only for strictly positive eigenvalues, in index order."""
    r = np.zeros(len(ev), dtype=float)
    for i in range(len(ev)):
        if ev[i] > 0:
            r[i] = draw(math.sqrt(ev[i]))
    return np.dot(np.asarray(evec), r)


def ph_covariance(i, gamma, gwl, T, phcut, dt, nmd, classical=False, zpmotion=True):
    """noise.py:74-79 -- A(w_i) = (dt nmd) equ(w_i) flinterp(w_i, gwl, gamma), hermitianised."""
    w = 2.0 * np.pi / dt / nmd * i
    return hermitianize(dt * nmd * equ(w, phcut, T, classical, zpmotion) * np.asarray(flinterp(w, gwl, gamma)))


def e_coefficients(i, bias, T, ecut, dt, nmd, classical=False, zpmotion=True):
    """noise.py:172-185 as three scalars: A = ce*efric + cp*exip + 1j*cm*exim."""
    w = 2.0 * np.pi / dt / nmd * i
    delta = dt * nmd
    aw = delta * equ(w, ecut, T, classical, zpmotion)
    awm = delta * equ(HBAR * w - bias, ecut, T, classical, zpmotion)
    awp = delta * equ(HBAR * w + bias, ecut, T, classical, zpmotion)
    return aw, (-aw + 0.5 * awm + 0.5 * awp), 0.5 * (awm - awp)


def e_covariance(i, efric, exim, exip, bias, T, ecut, dt, nmd, classical=False, zpmotion=True):
    """noise.py:172-186."""
    ce, cp, cm = e_coefficients(i, bias, T, ecut, dt, nmd, classical, zpmotion)
    return hermitianize(ce * np.asarray(efric) + cp * np.asarray(exip) + 1j * cm * np.asarray(exim))


def spectrum_to_series(x, dt, nmd):
    """noise.py:87-100 + functions.py:36-53: mirror to negative w, FFT, scale by 1/(dt nmd).

    x: [nmd/2+1, nc] positive-frequency samples.  Returns COMPLEX [nmd, nc]
    (callers take np.real, baths.py:191,408)."""
    hlen = nmd // 2
    neg = np.conjugate(x[hlen:0:-1, :])          # rows hlen, hlen-1, ..., 1
    full = np.concatenate((x[:hlen], neg), axis=0)
    if full.shape[0] != nmd:
        raise ValueError("MyFFT.iFourier1D: array length error!")
    return np.fft.fft(full, axis=0) * (2 * np.pi / dt / nmd) / 2 / np.pi


def noise_from_factors(L, xi, dt, nmd):
    """x_i = L_i xi_i per frequency, then spectrum_to_series; L: [nmd/2+1,nc,nc], xi: [nmd/2+1,nc]."""
    x = np.einsum("wij,wj->wi", L, xi)
    return np.real(spectrum_to_series(x, dt, nmd))


def eig_factor(A):
    """The reference's factor V sqrt(clamp+(lambda)) (noise.py:82-84 / 189-191 + vargau)."""
    ev, evec = np.linalg.eigh(A)
    return evec * np.sqrt(np.where(ev > 0, ev, 0.0))[None, :], ev


def phnoise(gamma, gwl, T, phcut, dt, nmd, draw, classical=False, zpmotion=True):
    """noise.py:50-100 (complex result; phbath.gnoi keeps np.real)."""
    hlen = nmd // 2
    x = []
    for i in range(hlen + 1):
        ev, evec = np.linalg.eigh(ph_covariance(i, gamma, gwl, T, phcut, dt, nmd, classical, zpmotion))
        x.append(vargau(ev, evec, draw))
    return spectrum_to_series(np.array(x), dt, nmd)


def enoise(efric, exim, exip, bias, T, ecut, dt, nmd, draw, classical=False, zpmotion=True):
    """noise.py:149-206 (complex result; ebath.gnoi keeps np.real)."""
    hlen = nmd // 2
    x = []
    for i in range(hlen + 1):
        ev, evec = np.linalg.eigh(e_covariance(i, efric, exim, exip, bias, T, ecut, dt, nmd, classical, zpmotion))
        x.append(vargau(ev, evec, draw))
    return spectrum_to_series(np.array(x), dt, nmd)


# --------------------------------------------------------------------------
# MD
# --------------------------------------------------------------------------
class Bath:
    """What md.force needs from a bath (baths.py:224-255, 448-458)."""

    def __init__(self, kind, cids, kernel, noise, dt, nmd, bias=0.0, exim=None, zeta1=None, zeta2=None):
        self.kind = kind                              # 'ph' | 'e'
        self.cids = np.asarray(cids, dtype=int)
        self.nc = len(self.cids)
        self.kernel = np.asarray(kernel, dtype=float)  # [ml, nc, nc]
        self.ml = self.kernel.shape[0]
        self.noise = np.asarray(noise, dtype=float)    # [nmd, nc]
        self.dt, self.nmd, self.bias = dt, nmd, bias
        z = np.zeros((self.nc, self.nc))
        self.exim = z if exim is None else np.asarray(exim, dtype=float)
        self.zeta1 = z if zeta1 is None else np.asarray(zeta1, dtype=float)
        self.zeta2 = z if zeta2 is None else np.asarray(zeta2, dtype=float)
        self.cur = np.zeros(nmd)

    @property
    def extra(self):
        """baths.py:233 -- the exim/zeta terms act only if ALL THREE have a non-zero entry."""
        return self.kind == "e" and bool(self.exim.any() and self.zeta1.any() and self.zeta2.any())

    def bforce(self, t, phis, qhis):
        f = self.noise[t % self.nmd]
        for i in range(self.ml):
            if self.ml == 1:
                f = f - self.kernel[i] @ phis[i][self.cids]
                if self.extra:
                    f = (f + (self.bias * self.exim) @ qhis[0][self.cids]
                         - (self.bias * self.zeta1) @ qhis[0][self.cids]
                         - (self.bias * self.zeta2) @ phis[0][self.cids])
            else:
                f = f - (self.kernel[i] @ phis[i][self.cids]) * self.dt
        out = np.zeros(len(phis[0]))
        out[self.cids] = f                            # noise.py:15-22 (mf)
        return out


def apply_constraint(f, constr):
    """md.py:782-794."""
    if constr is None:
        return f
    nf = np.array(f) * 1.0
    for c in constr:
        nf[np.asarray(list(c), dtype=int)] = 0
    return nf


def initialise(hw, Umat, T, constraint, rand):
    """md.py:308-326: random-phase normal-mode displacement / velocity; `rand()` stands for
    np.random.rand().  Returns (q, p)."""
    dis = np.zeros(len(hw))
    vel = np.zeros(len(hw))
    for i in range(len(hw)):
        if hw[i] < 0.01:
            am = 0.0
        else:
            am = ((bose(hw[i], T) + 0.5) * 2.0 / hw[i]) ** 0.5
        r = rand()
        dis = dis + Umat[:, i] * am * np.cos(2. * np.pi * r)
        vel = vel - hw[i] * Umat[:, i] * am * np.sin(2. * np.pi * r)
        dis = apply_constraint(dis, constraint)
        vel = apply_constraint(vel, constraint)
    return dis, vel


class LiteralMD:
    """Literal single-trajectory restatement of md.vv/force/potforce (md.py:367-474),
    including the physical history shift (functions.py:146-153) and the sameq
    force cache (md.py:449, 767-779).  This is the CPU baseline too."""

    def __init__(self, dyn, dt, nmd, baths, constraint=None, sameq_cache=True):
        self.dyn = np.asarray(dyn, dtype=float)
        self.nph = self.dyn.shape[0]
        self.dt, self.nmd = dt, nmd
        self.baths = baths
        self.ml = max([1] + [b.ml for b in baths])
        self.constraint = constraint
        self.sameq_cache = sameq_cache
        self.t = 0
        self.q = np.zeros(self.nph)
        self.p = np.zeros(self.nph)
        self.qhis = np.zeros((self.ml, self.nph))
        self.phis = np.zeros((self.ml, self.nph))
        self.etot = np.zeros(nmd)
        self.fbaths = [np.zeros(self.nph) for _ in baths]
        self.q0, self.f0 = None, None

    @staticmethod
    def _rpadleft(bs, b):
        if len(bs) > 1:
            return np.concatenate((np.array([b]), bs[:-1]), axis=0)
        return np.array([b])

    def potforce(self, q):
        if self.sameq_cache and self.q0 is not None and np.max(np.abs(q - self.q0)) < 10e-10:
            return self.f0
        f = -1.0 * (self.dyn @ q)
        self.q0, self.f0 = q, f
        return f

    def force(self, t, p, q, id=0):
        pf = self.potforce(q)
        if id == 0:
            tphis, tqhis = self.phis, self.qhis
        else:
            tphis, tqhis = self._rpadleft(self.phis, p), self._rpadleft(self.qhis, q)
        for i, b in enumerate(self.baths):
            self.fbaths[i] = b.bforce(t + id, tphis, tqhis)
            pf = pf + self.fbaths[i]
        return pf

    def vv(self):
        t, p, q, dt = int(self.t), self.p, self.q, self.dt
        self.etot[t % self.nmd] = 0.5 * np.dot(p, p)
        self.qhis = self._rpadleft(self.qhis, q)
        self.phis = self._rpadleft(self.phis, p)
        f = self.force(t, p, q, 0)
        pthalf = p + f * dt / 2.0
        qtt = q + p * dt + f * dt ** 2 / 2.0
        for i, b in enumerate(self.baths):
            b.cur[t % self.nmd] = np.dot(self.fbaths[i], p)
        f = self.force(t, pthalf, qtt, 1)
        ptt1 = pthalf + dt * f / 2.0
        f = self.force(t, ptt1, qtt, 1)
        ptt2 = pthalf + dt * f / 2.0
        ptt2 = apply_constraint(ptt2, self.constraint)
        qtt = apply_constraint(qtt, self.constraint)
        self.t, self.p, self.q, self.f = t + 1, ptt2, qtt, f


class EnsembleMD:
    """Single-tail restatement vectorised over trajectories.

    Per bath the friction tail  S = dt * sum_{j>=1} kernel[j] p_{t+1-j}[cids]  is
    evaluated ONCE per step and reused by force evaluations B, C and by A of the
    next step.  Histories are per-bath rings of p[cids] only.  Kernels may be
    full [ml,nc,nc] or diagonal [ml,nc].  Noise is [ntraj,nmd,nc].
    Equal to LiteralMD to ~1e-16 relative (tests/test_oracle_golden.py).
    """

    def __init__(self, dyn, dt, nmd, ntraj, constraint=None):
        self.K = np.asarray(dyn, dtype=float)
        self.nph = self.K.shape[0]
        self.dt, self.nmd, self.ntraj = dt, nmd, ntraj
        self.cons = None
        if constraint is not None:
            idx = []
            for c in constraint:
                idx.extend(list(c))
            self.cons = np.asarray(idx, dtype=int)
        self.baths = []
        self.t = 0
        self.q = np.zeros((ntraj, self.nph))
        self.p = np.zeros((ntraj, self.nph))
        self.etot = np.zeros((ntraj, nmd))
        self.Kq = None

    def add_bath(self, cids, kernel, noise, bias=0.0, exim=None, zeta1=None, zeta2=None, kind="ph"):
        kernel = np.asarray(kernel, dtype=float)
        cids = np.asarray(cids, dtype=int)
        nc = len(cids)
        b = dict(cids=cids, nc=nc, kernel=kernel, ml=kernel.shape[0], diag=(kernel.ndim == 2),
                 noise=np.asarray(noise, dtype=float), bias=bias,
                 ring=np.zeros((self.ntraj, kernel.shape[0], nc)),
                 tail=np.zeros((self.ntraj, nc)), cur=np.zeros((self.ntraj, self.nmd)), extra=False)
        if kind == "e" and exim is not None and zeta1 is not None and zeta2 is not None:
            if np.any(exim) and np.any(zeta1) and np.any(zeta2):
                b["extra"] = True
                b["Mq"] = bias * (np.asarray(exim) - np.asarray(zeta1))   # acts on q[cids]
                b["Mp"] = -bias * np.asarray(zeta2)                        # acts on p[cids]
        assert b["noise"].shape == (self.ntraj, self.nmd, nc)
        self.baths.append(b)
        return len(self.baths) - 1

    def _k0(self, b, x):
        c0 = self.dt if b["ml"] > 1 else 1.0
        if b["diag"]:
            return c0 * b["kernel"][0][None, :] * x
        return c0 * x @ b["kernel"][0].T

    def _bath_local(self, b, it, x, qc):
        f = b["noise"][:, it % self.nmd, :] - self._k0(b, x) - b["tail"]
        if b["extra"]:
            f = f + qc @ b["Mq"].T + x @ b["Mp"].T
        return f

    def _tail(self, b):
        """dt * sum_{j=1}^{ml-1} kernel[j] . ring[(head-j+1) mod ml]   (ring already holds p_t at head)."""
        ml = b["ml"]
        if ml == 1:
            return np.zeros((self.ntraj, b["nc"]))
        head = self.t % ml                                 # slot of p_t
        slots = (head - np.arange(0, ml - 1)) % ml          # j-1 = 0..ml-2  -> p_{t}, p_{t-1}, ...
        hist = b["ring"][:, slots, :]                       # [ntraj, ml-1, nc]
        if b["diag"]:
            return self.dt * np.einsum("jc,tjc->tc", b["kernel"][1:], hist)
        if "kmat" not in b:                                 # [(ml-1) nc, nc]: sum_j kernel[j] . p as ONE matrix product (BLAS)
            b["kmat"] = np.ascontiguousarray(b["kernel"][1:].transpose(0, 2, 1)).reshape((ml - 1) * b["nc"], b["nc"])
        return self.dt * (hist.reshape(self.ntraj, (ml - 1) * b["nc"]) @ b["kmat"])

    def step(self):
        dt, t = self.dt, self.t
        q, p = self.q, self.p
        self.etot[:, t % self.nmd] = 0.5 * np.einsum("ti,ti->t", p, p)
        if self.Kq is None:
            self.Kq = q @ self.K.T
        f = -self.Kq
        fa = []
        for b in self.baths:
            pc = p[:, b["cids"]]
            b["ring"][:, t % b["ml"], :] = pc               # push p_t
            fb = self._bath_local(b, t, pc, q[:, b["cids"]])
            fa.append(fb)
            f = f.copy()
            f[:, b["cids"]] += fb
        ph = p + f * dt / 2.0
        qn = q + p * dt + f * dt ** 2 / 2.0
        for b, fb in zip(self.baths, fa):
            b["cur"][:, t % self.nmd] = np.einsum("tc,tc->t", fb, p[:, b["cids"]])
            b["tail"] = self._tail(b)
        Kqn = qn @ self.K.T
        x = ph
        for _ in range(2):
            f = -Kqn
            for b in self.baths:
                f = f.copy()
                f[:, b["cids"]] += self._bath_local(b, t + 1, x[:, b["cids"]], qn[:, b["cids"]])
            x = ph + dt * f / 2.0
        if self.cons is not None and len(self.cons):
            x[:, self.cons] = 0.0
            qn[:, self.cons] = 0.0
            self.Kq = None
        else:
            self.Kq = Kqn
        self.p, self.q, self.t = x, qn, t + 1

    def run(self, n):
        for _ in range(n):
            self.step()


class ModalMD:
    """The same integrator (md.vv, md.py:367-411) propagated in the eigenbasis md.setDyn computes (md.py:266-281):
    K = U diag(lam) U^T, Q = U^T q, Pi = U^T p.  The harmonic force is diagonal there; only the bath dofs need real space.
    Prototype / specification of the CUDA engine's modal mode (sclmd_md_set_modes); must equal EnsembleMD to rounding.

    Restrictions (the engine falls back to real space otherwise): no constraints, baths on DISJOINT dof sets, every bath
    force diagonal in the momentum (diagonal kernels, no exim/zeta terms).

    With E_b = U[cids_b, :]  (p[cids_b] = E_b Pi, scatter U^T S_b f = E_b^T f) and the bath force fA_b(t) of evaluation A:
        Pt_t     = Pi_t + h/2 sum_b E_b^T fA_b(t)                          (momentum half-kicked by the baths)
        Q_{t+1}  = Q_t + h Pt_t - h^2/2 lam Q_t                            (md.py:392)
        on the bath dofs, in real space, with g = (K q)[cids] = E_b (lam Q):   evaluations A, B, C exactly as md.py:390-404
        Pt_{t+1} = Pt_t - h/2 lam (Q_t + Q_{t+1}) + h/2 sum_b E_b^T (fC_b(t) + fA_b(t+1))
        etot     = Pi.Pi/2,  Pi.Pi = Pt.Pt - h sum_b p_c.fA_b - h^2/4 sum_b |fA_b|^2     (E_b E_b'^T = delta_bb')
    Per step: ONE gather product [ntraj x nph].[nph x sum nc] and ONE scatter product [ntraj x sum nc].[sum nc x nph]
    (4 nph sum_nc flops per trajectory) instead of K.q (2 nph^2)."""

    def __init__(self, lam, U, dt, nmd, ntraj):
        self.lam, self.U = np.asarray(lam, dtype=float), np.asarray(U, dtype=float)
        self.nph = self.U.shape[0]
        self.dt, self.nmd, self.ntraj = dt, nmd, ntraj
        self.baths = []
        self.t = 0
        self.etot = np.zeros((ntraj, nmd))
        self.Q = np.zeros((ntraj, self.nph))
        self.Pt = np.zeros((ntraj, self.nph))
        self._primed = False

    def add_bath(self, cids, kernel, noise):
        kernel = np.asarray(kernel, dtype=float)
        assert kernel.ndim == 2, "modal mode: diagonal kernels only"
        cids = np.asarray(cids, dtype=int)
        for b in self.baths:
            assert not set(b["cids"]) & set(cids), "modal mode: baths must act on disjoint dofs"
        E = self.U[cids, :]
        b = dict(cids=cids, nc=len(cids), kernel=kernel, ml=kernel.shape[0], noise=np.asarray(noise, dtype=float), E=E, EL=E * self.lam[None, :],
                 ring=np.zeros((self.ntraj, kernel.shape[0], len(cids))), tail=np.zeros((self.ntraj, len(cids))),
                 cur=np.zeros((self.ntraj, self.nmd)))
        self.baths.append(b)

    def set_state(self, q, p):
        """real-space state in; the half-kicked momentum is formed when the first step starts"""
        self.Q = np.asarray(q, dtype=float) @ self.U
        self.Pi0 = np.asarray(p, dtype=float) @ self.U
        for b in self.baths:
            b["pc"] = np.asarray(p, dtype=float)[:, b["cids"]].copy()
        self._primed = False

    def _fb(self, b, it, x):
        c0 = self.dt if b["ml"] > 1 else 1.0
        return b["noise"][:, it % self.nmd, :] - c0 * b["kernel"][0][None, :] * x - b["tail"]

    def _tail(self, b):
        ml = b["ml"]
        if ml == 1:
            return np.zeros((self.ntraj, b["nc"]))
        head = self.t % ml
        slots = (head - np.arange(0, ml - 1)) % ml
        return self.dt * np.einsum("jc,tjc->tc", b["kernel"][1:], b["ring"][:, slots, :])

    def _observe(self):
        """etot, cur, ring push of evaluation A at the current time; needs b["fA"], b["pc"]"""
        t, h = self.t, self.dt
        pp = np.einsum("ti,ti->t", self.Pt, self.Pt)
        for b in self.baths:
            pp = pp - h * np.einsum("tc,tc->t", b["pc"], b["fA"]) - h * h / 4.0 * np.einsum("tc,tc->t", b["fA"], b["fA"])
            b["cur"][:, t % self.nmd] = np.einsum("tc,tc->t", b["fA"], b["pc"])
            b["ring"][:, t % b["ml"], :] = b["pc"]
        self.etot[:, t % self.nmd] = 0.5 * pp

    def _prime(self):
        h = self.dt
        self.Pt = self.Pi0.copy()
        for b in self.baths:
            b["g"] = self.Q @ b["EL"].T                     # (K q_t)[cids]
            b["fA"] = self._fb(b, self.t, b["pc"])
            self.Pt += h / 2.0 * (b["fA"] @ b["E"])
        self._observe()
        self._primed = True

    def step(self):
        if not self._primed:
            self._prime()
        h, t = self.dt, self.t
        Qn = self.Q + h * self.Pt - h * h / 2.0 * self.lam[None, :] * self.Q
        W = np.zeros_like(self.Q)
        for b in self.baths:
            b["tail"] = self._tail(b)                        # S'(t): ring holds p_t
            gn = Qn @ b["EL"].T                              # gather: (K q')[cids]
            ph = b["pc"] + h / 2.0 * (-b["g"] + b["fA"])
            x = ph
            for _ in range(2):                               # evaluations B and C (md.py:401-404)
                fb = self._fb(b, t + 1, x)
                x = ph + h / 2.0 * (-gn + fb)
            fC = fb
            b["pc"], b["g"] = x, gn
            b["fA"] = self._fb(b, t + 1, x)                  # evaluation A of step t+1: same noise row, same tail
            W += (fC + b["fA"]) @ b["E"]                     # scatter
        self.Pt = self.Pt - h / 2.0 * self.lam[None, :] * (self.Q + Qn) + h / 2.0 * W
        self.Q = Qn
        self.t = t + 1
        self._observe()

    def run(self, n):
        for _ in range(n):
            self.step()

    def state(self):
        """real-space (q, p) at the current time"""
        if not self._primed:
            return self.Q @ self.U.T, self.Pi0 @ self.U.T
        Pi = self.Pt.copy()
        for b in self.baths:
            Pi -= self.dt / 2.0 * (b["fA"] @ b["E"])
        return self.Q @ self.U.T, Pi @ self.U.T


# --------------------------------------------------------------------------
# NEGF (negf.py) and surface self-energy (selfenergy.py)
# --------------------------------------------------------------------------
def bosedist(omega, T):
    """negf.py:217-226."""
    if abs(T) < 1e-30:
        return 1 / (np.exp(RPC * omega * np.iinfo(np.int32).max) - 1)
    elif abs(omega / T) < 1e-30:
        return np.iinfo(np.int32).max
    return 1 / (np.exp(RPC * omega / BC / T) - 1)


def bpt_reduce_index(dofs, nfixed0):
    """negf.py:195-204: bath dofs are given in the unreduced 3N numbering; the
    leading fixed block is removed so reduced index = dof - len(fixed[0])."""
    return np.asarray(list(dofs), dtype=int) - nfixed0


def bpt_retargf(K, omega, damp, idxL, idxR):
    """negf.py:206-208 with negf.py:153-157 folded in (Sigma = -i w/damp on bath dofs)."""
    n = K.shape[0]
    A = (omega + 1e-9j) ** 2 * np.identity(n) - K
    d = np.zeros(n, dtype=complex)
    d[idxL] += -1j * omega / damp
    d[idxR] += -1j * omega / damp
    return np.linalg.inv(A - np.diag(d))


def bpt_tm(K, omega, damp, idxL, idxR):
    """negf.py:240-242: Re Tr[G Gamma_L G^dagger Gamma_R], Gamma = -i(Sigma - Sigma^dagger) = (2w/damp) on bath dofs... sign: -i(-iw/d - (+iw/d)) = -2w/d."""
    G = bpt_retargf(K, omega, damp, idxL, idxR)
    n = K.shape[0]
    gl = np.zeros(n)
    gr = np.zeros(n)
    gl[idxL] = -2.0 * omega / damp
    gr[idxR] = -2.0 * omega / damp
    return float(np.real(np.trace(G @ np.diag(gl) @ G.conj().T @ np.diag(gr))))


def bpt_ps_nobias(K, omega, T, damp, idxL, idxR, sel):
    """negf.py:232: -2 w^2 n_B(w,T) Tr Im G[sel,sel]."""
    G = bpt_retargf(K, omega, damp, idxL, idxR)
    return float(-2 * omega ** 2 * bosedist(omega, T) * np.trace(np.imag(G[sel][:, sel])))


def bpt_bias_matrices(K, omega, T, damp, idxL, idxR, b0, bdamp, chiplus, chiminus, bias):
    """negf.py:153-193, 206-212 with a biased block on dofs [b0, b0+nb): returns (G^r, Sigma^K_total, G^a).
    `bias` is in angular units (bias_eV / rpc), as bpt.setbias stores it."""
    n = K.shape[0]
    nb = len(bdamp)
    sl = np.zeros((n, n), dtype=complex)
    sr = np.zeros((n, n), dtype=complex)
    sl[idxL, idxL] = -1j * omega / damp
    sr[idxR, idxR] = -1j * omega / damp
    sb = np.zeros((n, n), dtype=complex)
    blk = slice(b0, b0 + nb)
    sb[blk, blk] = -1j * omega * np.asarray(bdamp) - bias * np.asarray(chiminus)
    z2 = (omega + 1e-9j) ** 2 * np.identity(n)
    Gr = np.linalg.inv(z2 - K - sl - sr - sb)
    Ga = np.linalg.inv(z2 - K - sl.conj().T - sr.conj().T - sb.conj().T)
    with np.errstate(all="ignore"):
        n0 = bosedist(omega, T)
        semat = np.zeros((n, n), dtype=complex)
        semat[blk, blk] = ((np.asarray(chiplus) - 1j * np.asarray(chiminus)) * (omega + bias) * (2 * bosedist(omega + bias, T) - 2 * n0)
                           + (np.asarray(chiplus) + 1j * np.asarray(chiminus)) * (omega - bias) * (2 * bosedist(omega - bias, T) - 2 * n0)) / 2
        sk = -2 * np.imag(sl) * n0 + -2 * np.imag(sr) * n0 + (1j * sb) * 2 * n0 + semat
    return Gr, sk, Ga


def bpt_ps_bias(K, omega, T, damp, idxL, idxR, b0, bdamp, chiplus, chiminus, bias, sel):
    """negf.py:236."""
    Gr, sk, Ga = bpt_bias_matrices(K, omega, T, damp, idxL, idxR, b0, bdamp, chiplus, chiminus, bias)
    return float(omega ** 2 * np.trace(np.real((Gr @ sk @ Ga)[sel][:, sel])))


def bpt_tm_bias(K, omega, T, damp, idxL, idxR, b0, bdamp, chiplus, chiminus, bias):
    """negf.py:240-242 with retargf including the bias self-energy."""
    Gr, _, _ = bpt_bias_matrices(K, omega, T, damp, idxL, idxR, b0, bdamp, chiplus, chiminus, bias)
    n = K.shape[0]
    gl = np.zeros(n)
    gr = np.zeros(n)
    gl[idxL] = -2.0 * omega / damp
    gr[idxR] = -2.0 * omega / damp
    return float(np.real(np.trace(Gr @ np.diag(gl) @ Gr.conj().T @ np.diag(gr))))


def thermalcurrent(tmnumber, T, delta):
    """negf.py:245-270 trapezoid, nW."""
    n = len(tmnumber) - 1
    arr = np.array([RPC * w / 2 / np.pi * tm * (bosedist(w, T * (1 + 0.5 * delta)) - bosedist(w, T * (1 - 0.5 * delta)))
                    for w, tm in tmnumber])
    return (float(tmnumber[-1, 0] - tmnumber[0, 0]) / n / 2.) * (2 * arr.sum() - arr[0] - arr[-1]) * 1.60217662 * 1e2


def sig_sgf(K00, K11, K01, K10, omega, eta, direction):
    """selfenergy.py:105-131 (Sancho-Rubio; beta = alpha^T re-derived every iteration)."""
    if direction == "R":
        s, e, alpha = K00.astype(complex), K11.astype(complex), K01.astype(complex)
    elif direction == "L":
        s, e, alpha = K11.astype(complex), K00.astype(complex), K10.astype(complex)
    else:
        raise ValueError("Wrong direction, should only be R or L")
    it = 0
    z = (omega + eta * 1j) ** 2
    while np.linalg.norm(alpha) > 1e-8:
        g = np.linalg.inv(z * np.identity(len(e)) - e)
        beta = alpha.T
        agb = alpha @ g @ beta
        s = s + agb
        e = e + agb + beta @ g @ alpha
        alpha = alpha @ g @ alpha
        it += 1
        if it >= 100:
            raise ValueError("Iteration number exceeded 100, please increase eta")
    return np.linalg.inv(z * np.identity(len(s)) - s), it


def sig_selfenergy(K00, K11, K01, K10, omega, eta, direction):
    """selfenergy.py:133-140."""
    g, _ = sig_sgf(K00, K11, K01, K10, omega, eta, direction)
    if direction == "R":
        return K01 @ g @ K10
    return K10 @ g @ K01


def sig_tm(K00, K11, K01, K10, omega, eta):
    """selfenergy.py:145-151."""
    sl = sig_selfenergy(K00, K11, K01, K10, omega, eta, "L")
    sr = sig_selfenergy(K00, K11, K01, K10, omega, eta, "R")
    G = np.linalg.inv((omega + 1e-8 * 1j) ** 2 * np.identity(len(K00)) - K00 - sl - sr)
    gam = lambda P: -1j * (P - P.conj().T)
    return float(np.real(np.trace(G @ gam(sl) @ G.conj().T @ gam(sr))))


def blocked_tail_parts(kernel, ring, t0, s, dt, tb=32, sr=8, seg=None):
    """The friction tail S'(t0 + s) of a diagonal kernel as the device evaluates it in time-blocked mode (test infrastructure: the
    specification of k_tail_near / k_tail_far_mma).  ring[slot] = p_{t'} with t' mod ml == slot, rows up to
    t0 + s present; kernel [ml, nc] (zero beyond row ml - 1).  Returns (near, mid, farfar):
      near    = dt sum_{j=1}^{s+1} k[j] p_{t0+s+1-j}                  rows of the current block
      mid     = dt sum_{d=0}^{tb-1} k[s+2+d] p_{t0-1-d}               rows of the previous block (one short pass at the block boundary)
      farfar  = dt sum_{d>=tb}      k[s+2+d] p_{t0-1-d}               older rows: final one block early, worked off one slice per step;
                                                                      `seg` = list of (d_lo, d_hi) age ranges summed separately and then
                                                                      in order (the partial slots of the device), default one range
    near + mid + farfar == dt sum_{j=1}^{ml-1} k[j] p_{t0+s+1-j} (md.py:386-387 / baths.py:453-457 restricted to j >= 1)."""
    ml, nc = kernel.shape
    kp = np.vstack([kernel, np.zeros((tb + 2 + sr, nc))])      # zero rows past ml - 1 retire ring slots that were overwritten
    t = t0 + s
    near = np.zeros(nc)
    for j in range(1, s + 2):
        near += kp[j] * ring[(t + 1 - j) % ml]
    def ages(lo, hi):
        acc = np.zeros(nc)
        for d in range(lo, hi):
            acc += kp[s + 2 + d] * ring[(t0 - 1 - d) % ml]
        return acc
    mid = ages(0, tb)
    seg = seg or [(tb, ml)]
    farfar = np.zeros(nc)
    for lo, hi in seg:
        farfar = farfar + dt * ages(lo, hi)
    return dt * near, dt * mid, farfar


def hankel_tile(kvec, x, n):
    """fragment F(x) of the Hankel operand of the tensor-pipe far pass: an 8 x 4 tile H[s][a] = k[x + 2 + s + a] of step tile n at ages
    x - 8 n .. x - 8 n + 3 (k_tail_far_mma: one new fragment per k-step and dof, a tile of step tile n is F(d0 + 8 n))"""
    return np.array([[kvec[x + 2 + s + a] for a in range(4)] for s in range(8)])
