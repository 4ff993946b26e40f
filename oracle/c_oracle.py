"""ctypes wrapper of oracle/nde_oracle.c (TEST INFRASTRUCTURE; see that file's header)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "liboracle.so")


class OrcParams(C.Structure):
    _fields_ = [("nz", C.c_int32), ("H", C.c_float), ("tau", C.c_float), ("f", C.c_float), ("g", C.c_float),
                ("alpha", C.c_float), ("nu0", C.c_float), ("nu_m", C.c_float), ("Ric", C.c_float), ("dRi", C.c_float),
                ("Pr", C.c_float), ("kappa", C.c_float), ("eps", C.c_float), ("mu", C.c_float * 6),
                ("sigma", C.c_float * 6), ("flags", C.c_uint32), ("diurnal_period", C.c_float),
                ("reserved", C.c_float * 10)]


def build(force=False):
    src = os.path.join(HERE, "nde_oracle.c")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-shared", "-o", SO, src, "-lm"])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        _lib = C.CDLL(SO)
        _lib.orc_rollout.argtypes = [C.POINTER(OrcParams), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p,
                                     C.c_int]
        _lib.orc_rollout.restype = C.c_int
        _lib.orc_rollout_f64.argtypes = _lib.orc_rollout.argtypes
        _lib.orc_rollout_f64.restype = C.c_int
        _lib.orc_max_threads.restype = C.c_int
        _lib.orc_num_procs.restype = C.c_int
    return _lib


def params_struct(p) -> OrcParams:
    q = OrcParams()
    q.nz = p.Nz
    q.H, q.tau, q.f, q.g, q.alpha = p.H, p.tau, p.f, p.g, p.alpha
    q.nu0, q.nu_m, q.Ric, q.dRi, q.Pr, q.kappa, q.eps = p.nu0, p.nu_m, p.Ric, p.dRi, p.Pr, p.kappa, p.eps
    for i in range(6):
        q.mu[i], q.sigma[i] = p.mu[i], p.sigma[i]
    q.flags = p.flags
    q.diurnal_period = p.diurnal_period
    return q


def rollout(x0, bc, model, p, scheme, dt_nd, n_steps, save_every, t0=0.0, n_threads=0, f64=False):
    """Float32 solve (f64=False) or the same Float32-specified problem solved in double (f64=True: the noise floor)."""
    x0 = np.ascontiguousarray(x0, dtype=np.float32)
    bc = np.ascontiguousarray(bc, dtype=np.float32)
    w = np.ascontiguousarray(model.flat(), dtype=np.float32)
    B = x0.shape[0]
    n_save = n_steps // save_every + 1 if save_every > 0 else 1
    out = np.empty((B, n_save, x0.shape[1]), dtype=np.float64 if f64 else np.float32)
    sizes = (C.c_int * len(model.sizes))(*model.sizes)
    q = params_struct(p)
    fn = lib().orc_rollout_f64 if f64 else lib().orc_rollout
    rc = fn(C.byref(q), sizes, len(model.sizes), model.act, w.ctypes.data, x0.ctypes.data, bc.ctypes.data,
                           B, scheme, dt_nd, t0, n_steps, save_every, out.ctypes.data, n_threads)
    if rc != 0:
        raise RuntimeError("orc_rollout failed")
    return out


def max_threads() -> int:
    return int(lib().orc_max_threads())


def num_procs() -> int:
    """Host cores, independent of OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    return int(lib().orc_num_procs())
