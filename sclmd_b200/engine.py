"""Thin object wrapper over the sclmd_md_* C ABI (include/sclmd_b200.h).

`MDEngine` is what sclmd_b200.md.md drives; tests and bench.py use it directly
when they need raw access (host buffers in, host buffers out)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import as_f64, as_i32, check, dptr, iptr

KERNEL_FULL, KERNEL_DIAG = 0, 1


class MDEngine:
    def __init__(self, nph, ntraj, dt, nmd, device=0):
        self.nph, self.ntraj, self.dt, self.nmd, self.device = int(nph), int(ntraj), float(dt), int(nmd), int(device)
        self._h = C.c_void_p()
        self._baths = []  # (nc, ml)
        check(_lib.lib().sclmd_md_create(self.nph, self.ntraj, self.dt, self.nmd, self.device, C.byref(self._h)))

    def close(self):
        if self._h:
            _lib.lib().sclmd_md_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- problem definition
    def set_dyn(self, K):
        K = as_f64(K, (self.nph, self.nph))
        check(_lib.lib().sclmd_md_set_dyn(self._h, dptr(K)))

    def set_modes(self, lam, U):
        """eigen-decomposition of the matrix given to set_dyn (K = U diag(lam) U^T, eigenvectors as columns, md.py:266-281):
        enables the eigenbasis propagation where the problem allows it"""
        lam = as_f64(lam, (self.nph,))
        U = as_f64(U, (self.nph, self.nph))
        check(_lib.lib().sclmd_md_set_modes(self._h, dptr(lam), dptr(U)))

    def set_modal(self, on=True):
        check(_lib.lib().sclmd_md_set_modal(self._h, 1 if on else 0))

    def modal_active(self):
        return bool(check(_lib.lib().sclmd_md_modal_active(self._h)))

    def set_constraint(self, idx):
        idx = as_i32(idx)
        check(_lib.lib().sclmd_md_set_constraint(self._h, iptr(idx), len(idx)))

    def add_bath(self, cids, kernel, Mq=None, Mp=None):
        """kernel: [ml,nc,nc] (full) or [ml,nc] (diagonal)."""
        cids = as_i32(cids)
        nc = len(cids)
        kernel = as_f64(kernel)
        if kernel.ndim == 3 and kernel.shape[1:] == (nc, nc):
            kind = KERNEL_FULL
        elif kernel.ndim == 2 and kernel.shape[1] == nc:
            kind = KERNEL_DIAG
        else:
            raise ValueError("kernel must be [ml,%d,%d] or [ml,%d], got %s" % (nc, nc, nc, kernel.shape))
        ml = kernel.shape[0]
        Mq = None if Mq is None else as_f64(Mq, (nc, nc))
        Mp = None if Mp is None else as_f64(Mp, (nc, nc))
        out = C.c_int32(-1)
        check(_lib.lib().sclmd_md_add_bath(self._h, iptr(cids), nc, ml, dptr(kernel), kind, dptr(Mq), dptr(Mp), C.byref(out)))
        self._baths.append((nc, ml))
        return out.value

    # ---- data movement
    def set_noise(self, bath, noise, traj0=0):
        nc, _ = self._baths[bath]
        noise = as_f64(noise)
        if noise.ndim == 2:
            noise = noise[None]
        if noise.shape[1:] != (self.nmd, nc):
            raise ValueError("noise must be [ntraj,%d,%d], got %s" % (self.nmd, nc, noise.shape))
        noise = np.ascontiguousarray(noise)
        check(_lib.lib().sclmd_md_set_noise(self._h, bath, traj0, noise.shape[0], dptr(noise)))

    def get_noise(self, bath, traj0=0, ntraj=None):
        nc, _ = self._baths[bath]
        n = self.ntraj - traj0 if ntraj is None else ntraj
        out = np.empty((n, self.nmd, nc))
        check(_lib.lib().sclmd_md_get_noise(self._h, bath, traj0, n, dptr(out)))
        return out

    def set_noise_rows(self, bath, slab0, rows):
        """rows: [nslab, ntraj, nc] time-major (may live in pinned host memory)."""
        nc, _ = self._baths[bath]
        if rows.dtype != np.float64 or not rows.flags.c_contiguous or rows.shape[1:] != (self.ntraj, nc):
            raise ValueError("rows must be C-contiguous float64 [nslab,%d,%d]" % (self.ntraj, nc))
        check(_lib.lib().sclmd_md_set_noise_rows(self._h, bath, int(slab0), rows.shape[0], dptr(rows)))

    def step_observables(self, slab, out=None):
        if out is None:
            out = np.empty((1 + len(self._baths), self.ntraj))
        check(_lib.lib().sclmd_md_get_step_observables(self._h, int(slab), dptr(out)))
        return out

    def set_persistent(self, on=True):
        check(_lib.lib().sclmd_md_set_persistent(self._h, 1 if on else 0))

    def set_overlap(self, on=True):
        check(_lib.lib().sclmd_md_set_overlap(self._h, 1 if on else 0))

    def set_tail_block(self, on=True):
        check(_lib.lib().sclmd_md_set_tail_block(self._h, int(on)))

    def profile_all(self):
        ms = np.zeros(4)
        n = np.zeros(4, dtype=np.int64)
        check(_lib.lib().sclmd_md_get_profile_all(self._h, dptr(ms), n.ctypes.data_as(_lib.c_int64_p)))
        names = ("tail_direct", "potforce", "tail_far", "tail_near")
        return {k: dict(ms=float(ms[i]), launches=int(n[i])) for i, k in enumerate(names)}

    def profile_ex(self):
        ms = np.zeros(8)
        n = np.zeros(8, dtype=np.int64)
        check(_lib.lib().sclmd_md_get_profile_ex(self._h, dptr(ms), n.ctypes.data_as(_lib.c_int64_p)))
        names = ("tail_direct", "potforce", "tail_far", "tail_near", "modal_scatter", "modal_gather", "modal_bath", "modal_update")
        return {k: dict(ms=float(ms[i]), launches=int(n[i])) for i, k in enumerate(names)}

    def set_profiling(self, on=True):
        check(_lib.lib().sclmd_md_set_profiling(self._h, 1 if on else 0))

    def profile(self):
        a, b = C.c_double(0), C.c_double(0)
        na, nb = C.c_int64(0), C.c_int64(0)
        check(_lib.lib().sclmd_md_get_profile(self._h, C.byref(a), C.byref(na), C.byref(b), C.byref(nb)))
        return dict(tail_ms=a.value, tail_launches=na.value, potforce_ms=b.value, potforce_launches=nb.value)

    def set_state(self, q=None, p=None, t=-1):
        q = None if q is None else as_f64(np.broadcast_to(q, (self.ntraj, self.nph)))
        p = None if p is None else as_f64(np.broadcast_to(p, (self.ntraj, self.nph)))
        check(_lib.lib().sclmd_md_set_state(self._h, dptr(q), dptr(p), int(t)))

    def get_state(self):
        q = np.empty((self.ntraj, self.nph))
        p = np.empty((self.ntraj, self.nph))
        t = C.c_int64(0)
        check(_lib.lib().sclmd_md_get_state(self._h, dptr(q), dptr(p), C.byref(t)))
        return q, p, t.value

    def reset_history(self):
        check(_lib.lib().sclmd_md_reset_history(self._h))

    def get_history(self, bath):
        nc, ml = self._baths[bath]
        out = np.empty((self.ntraj, ml, nc))
        check(_lib.lib().sclmd_md_get_history(self._h, bath, dptr(out)))
        return out

    def set_history(self, bath, phis):
        nc, ml = self._baths[bath]
        phis = as_f64(phis, (self.ntraj, ml, nc))
        check(_lib.lib().sclmd_md_set_history(self._h, bath, dptr(phis)))

    # ---- stepping / observables
    def run(self, nsteps):
        """advance every trajectory by nsteps; returns device milliseconds (CUDA events)."""
        ms = C.c_float(0)
        check(_lib.lib().sclmd_md_run(self._h, int(nsteps), C.byref(ms)))
        return ms.value

    def run_async(self, nsteps):
        """enqueue nsteps without waiting; any getter synchronises"""
        check(_lib.lib().sclmd_md_run(self._h, int(nsteps), None))

    # ---- md.f / md.fbaths of the last step
    def set_force_output(self, on=True):
        check(_lib.lib().sclmd_md_set_force_output(self._h, 1 if on else 0))

    def last_force(self):
        out = np.empty((self.ntraj, self.nph))
        check(_lib.lib().sclmd_md_get_force(self._h, dptr(out)))
        return out

    def last_bath_force(self, bath, evaluation=0):
        """bath force of the last step at evaluation A (0) or C (2): [ntraj, nc]"""
        out = np.empty((self.ntraj, self._baths[bath][0]))
        check(_lib.lib().sclmd_md_get_bath_force(self._h, bath, evaluation, dptr(out)))
        return out

    # ---- force drivers (md.AddPotential): potential force from a host callback, everything else on the device
    def set_external_force(self, on=True):
        check(_lib.lib().sclmd_md_set_external_force(self._h, 1 if on else 0))

    def step_with_driver(self, force, q_now=None):
        """one step; `force(q[ntraj, nph]) -> f[ntraj, nph]` is the driver.  q_now: current positions (fetched if the force
        at them is needed and they are not given)"""
        L = _lib.lib()
        need = L.sclmd_md_force_needed(self._h)
        check(min(need, 0))
        if need:
            if q_now is None:
                q_now = self.get_state()[0]
            f0 = as_f64(force(np.asarray(q_now, dtype=float).reshape(self.ntraj, self.nph)), (self.ntraj, self.nph))
            check(L.sclmd_md_set_force(self._h, dptr(f0)))
        qt = np.empty((self.ntraj, self.nph))
        check(L.sclmd_md_step_begin(self._h, dptr(qt)))
        f1 = as_f64(force(qt), (self.ntraj, self.nph))
        check(L.sclmd_md_step_end(self._h, dptr(f1)))

    def current(self, bath):
        out = np.empty((self.ntraj, self.nmd))
        check(_lib.lib().sclmd_md_get_current(self._h, bath, dptr(out)))
        return out

    def set_current(self, bath, cur):
        cur = as_f64(np.asarray(cur, dtype=float).reshape(self.ntraj, self.nmd))
        check(_lib.lib().sclmd_md_set_current(self._h, bath, dptr(cur)))

    def set_etot(self, etot):
        etot = as_f64(np.asarray(etot, dtype=float).reshape(self.ntraj, self.nmd))
        check(_lib.lib().sclmd_md_set_etot(self._h, dptr(etot)))

    def etot(self):
        out = np.empty((self.ntraj, self.nmd))
        check(_lib.lib().sclmd_md_get_etot(self._h, dptr(out)))
        return out

    def current_sums(self, bath):
        out = np.empty(self.ntraj)
        check(_lib.lib().sclmd_md_get_current_sums(self._h, bath, dptr(out)))
        return out

    def launch_count(self):
        return int(_lib.lib().sclmd_md_launch_count(self._h))

    def time_tail(self, bath, reps=5):
        ms = C.c_float(0)
        check(_lib.lib().sclmd_md_time_tail(self._h, bath, reps, C.byref(ms)))
        return ms.value

    def time_potforce(self, reps=5):
        ms = C.c_float(0)
        check(_lib.lib().sclmd_md_time_potforce(self._h, reps, C.byref(ms)))
        return ms.value
