"""ctypes binding of libsclmd_b200.so (include/sclmd_b200.h).

There is no CPU fallback: if the library is missing or no CUDA device is
visible, every compute call raises.  Nothing in this package imports the CPU
oracle."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(HERE, "libsclmd_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
c_float_p = C.POINTER(C.c_float)


class SclmdError(RuntimeError):
    pass


class NoConvergence(SclmdError, ValueError):
    """sig.sgf exceeded 100 decimation iterations (selfenergy.py:127-130 raises ValueError)."""


_PROTOS = {
    "sclmd_last_error": (C.c_char_p, []),
    "sclmd_version": (C.c_int, []),
    "sclmd_device_count": (C.c_int, []),
    "sclmd_device_info": (C.c_int, [C.c_int, c_int32_p, c_int32_p, c_int32_p, C.POINTER(C.c_uint64)]),
    "sclmd_probe_fp64": (C.c_int, [C.c_int, C.c_int, c_double_p]),
    "sclmd_md_create": (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "sclmd_md_destroy": (C.c_int, [C.c_void_p]),
    "sclmd_md_set_dyn": (C.c_int, [C.c_void_p, c_double_p]),
    "sclmd_md_set_constraint": (C.c_int, [C.c_void_p, c_int32_p, C.c_int]),
    "sclmd_md_set_modes": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    "sclmd_md_set_modal": (C.c_int, [C.c_void_p, C.c_int]),
    "sclmd_md_modal_active": (C.c_int, [C.c_void_p]),
    "sclmd_md_get_profile_ex": (C.c_int, [C.c_void_p, c_double_p, c_int64_p]),
    "sclmd_md_add_bath": (C.c_int, [C.c_void_p, c_int32_p, C.c_int, C.c_int, c_double_p, C.c_int, c_double_p, c_double_p, c_int32_p]),
    "sclmd_md_set_noise": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p]),
    "sclmd_md_get_noise": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p]),
    "sclmd_md_set_noise_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p]),
    "sclmd_md_get_step_observables": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "sclmd_md_set_overlap": (C.c_int, [C.c_void_p, C.c_int]),
    "sclmd_md_set_persistent": (C.c_int, [C.c_void_p, C.c_int]),
    "sclmd_md_set_force_output": (C.c_int, [C.c_void_p, C.c_int]),
    "sclmd_md_get_force": (C.c_int, [C.c_void_p, c_double_p]),
    "sclmd_md_get_bath_force": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p]),
    "sclmd_md_set_external_force": (C.c_int, [C.c_void_p, C.c_int]),
    "sclmd_md_force_needed": (C.c_int, [C.c_void_p]),
    "sclmd_md_set_force": (C.c_int, [C.c_void_p, c_double_p]),
    "sclmd_md_step_begin": (C.c_int, [C.c_void_p, c_double_p]),
    "sclmd_md_step_end": (C.c_int, [C.c_void_p, c_double_p]),
    "sclmd_md_set_tail_block": (C.c_int, [C.c_void_p, C.c_int]),
    "sclmd_md_get_profile_all": (C.c_int, [C.c_void_p, c_double_p, c_int64_p]),
    "sclmd_md_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "sclmd_md_get_profile": (C.c_int, [C.c_void_p, c_double_p, c_int64_p, c_double_p, c_int64_p]),
    "sclmd_md_set_state": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int64]),
    "sclmd_md_get_state": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_int64_p]),
    "sclmd_md_reset_history": (C.c_int, [C.c_void_p]),
    "sclmd_md_get_history": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "sclmd_md_set_history": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "sclmd_md_run": (C.c_int, [C.c_void_p, C.c_int64, c_float_p]),
    "sclmd_md_get_current": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "sclmd_md_get_etot": (C.c_int, [C.c_void_p, c_double_p]),
    "sclmd_md_set_current": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "sclmd_md_set_etot": (C.c_int, [C.c_void_p, c_double_p]),
    "sclmd_md_get_current_sums": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "sclmd_md_launch_count": (C.c_int64, [C.c_void_p]),
    "sclmd_dgemm_nt": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, C.c_double, c_double_p]),
    "sclmd_md_time_tail": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_float_p]),
    "sclmd_md_time_potforce": (C.c_int, [C.c_void_p, C.c_int, c_float_p]),
    "sclmd_noise_plan_create": (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, c_double_p, C.c_int, c_int32_p, c_double_p, c_double_p, C.POINTER(C.c_void_p)]),
    "sclmd_noise_plan_destroy": (C.c_int, [C.c_void_p]),
    "sclmd_noise_plan_dims": (C.c_int, [C.c_void_p, c_int32_p, c_int32_p]),
    "sclmd_noise_plan_is_complex": (C.c_int, [C.c_void_p]),
    "sclmd_noise_plan_get_factors": (C.c_int, [C.c_void_p, c_double_p]),
    "sclmd_noise_plan_set_factors": (C.c_int, [C.c_void_p, c_double_p, C.c_int]),
    "sclmd_noise_plan_generate": (C.c_int, [C.c_void_p, C.c_int, c_double_p, C.c_uint64, C.c_int64, c_double_p]),
    "sclmd_noise_plan_generate_into": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "sclmd_noise_plan_launch_count": (C.c_int64, [C.c_void_p]),
    "sclmd_noise_plan_get_profile": (C.c_int, [C.c_void_p, c_double_p]),
    "sclmd_md_generate_noise": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_int64]),
    "sclmd_fft": (C.c_int, [C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, C.c_int, C.c_double, c_double_p, c_double_p]),
    "sclmd_cos_transform": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, C.c_double, C.c_double, c_double_p]),
    "sclmd_gamt": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]),
    "sclmd_release_workspace": (C.c_int, []),
    "sclmd_bpt_create": (C.c_int, [C.c_int, C.c_int, c_double_p, c_int32_p, C.c_int, c_int32_p, C.c_int, C.c_double, C.POINTER(C.c_void_p)]),
    "sclmd_bpt_destroy": (C.c_int, [C.c_void_p]),
    "sclmd_bpt_set_bias": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, C.c_double]),
    "sclmd_bpt_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "sclmd_bpt_get_profile": (C.c_int, [C.c_void_p, c_double_p, c_int64_p, c_double_p, c_double_p]),
    "sclmd_bpt_tm": (C.c_int, [C.c_void_p, c_double_p, C.c_int, c_double_p]),
    "sclmd_bpt_ps": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int, c_int32_p, C.c_int, c_double_p]),
    "sclmd_bpt_green": (C.c_int, [C.c_void_p, c_double_p, C.c_int, C.c_int, c_double_p]),
    "sclmd_bpt_ps_bias": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, C.c_int, c_int32_p, C.c_int, c_double_p]),
    "sclmd_sig_selfenergy": (C.c_int, [C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p, C.c_double, C.c_char, c_double_p, C.c_int, c_double_p, c_int32_p]),
    "sclmd_sig_sgf": (C.c_int, [C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p, C.c_double, C.c_char, c_double_p, C.c_int, c_double_p, c_int32_p]),
    "sclmd_sig_green": (C.c_int, [C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p, C.c_double, c_double_p, C.c_int, c_double_p]),
    "sclmd_sig_tm": (C.c_int, [C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p, C.c_double, c_double_p, C.c_int, c_double_p]),
}

_lib = None


def exported_symbols():
    """Names declared in include/sclmd_b200.h that the binding expects."""
    return sorted(_PROTOS)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIBPATH):
            raise SclmdError("libsclmd_b200.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'`. "
                             "There is no CPU fallback." % LIBPATH)
        L = C.CDLL(LIBPATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc < 0:
        msg = lib().sclmd_last_error().decode("utf-8", "replace")
        if rc == -4:
            raise NoConvergence(msg)
        raise SclmdError(msg)
    return rc


def dptr(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def iptr(a):
    return None if a is None else a.ctypes.data_as(c_int32_p)


def as_f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError("expected array of shape %s, got %s" % (tuple(shape), tuple(a.shape)))
    return a


def as_i32(a):
    return np.ascontiguousarray(np.asarray(list(a) if not isinstance(a, np.ndarray) else a), dtype=np.int32)


def dgemm_nt(A, B, alpha=1.0, device=0):
    """alpha * A . B^T on the device (A [M,K], B [N,K] host arrays)"""
    A = np.ascontiguousarray(np.atleast_2d(A), dtype=np.float64)
    B = np.ascontiguousarray(np.atleast_2d(B), dtype=np.float64)
    if A.shape[1] != B.shape[1]:
        raise ValueError("dgemm_nt: inner dimensions differ")
    out = np.empty((A.shape[0], B.shape[0]))
    check(lib().sclmd_dgemm_nt(int(device), A.shape[0], B.shape[0], A.shape[1], dptr(A), dptr(B), float(alpha), dptr(out)))
    return out


def device_count():
    return check(lib().sclmd_device_count())
