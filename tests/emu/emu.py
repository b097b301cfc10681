"""Loader for the host lock-step build of the kernel source (tests only)."""
import ctypes as C
import os
import subprocess

import numpy as np

from metrotrpl_b200 import _capi

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libtrpl_emu.so")
SRC = os.path.join(HERE, "trpl_emu.cpp")
CSRC = os.path.join(HERE, "..", "..", "metrotrpl_b200", "csrc")


def build(force=False):
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return SO
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fopenmp", "-ffp-contract=off", "-shared",
                           "-fPIC", "-o", SO, SRC])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(SO)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        _lib.trpl_emu_loglik_batch.argtypes = [C.c_int32, C.c_int32, C.POINTER(_capi.MeasDesc), C.c_int32,
                                               dp, dp, dp, dp, C.c_int32, dp, dp,
                                               C.POINTER(_capi.SolverOpts), dp, ip, ip, dp, dp, dp, C.c_int32, dp]
    return _lib


def loglik_batch(prob, params, aux, opts, want_curves=True, ladder=None):
    params = np.ascontiguousarray(params, dtype=np.float64)
    aux = np.ascontiguousarray(aux, dtype=np.float64)
    n_sets = params.shape[0]
    logll = np.empty((n_sets, prob.n_meas, 3))
    status = np.empty((n_sets, prob.n_meas), dtype=np.int32)
    nsteps = np.empty((n_sets, prob.n_meas, 2), dtype=np.int32)
    curves = np.empty((n_sets, prob.n_times_total)) if want_curves else None
    p = _capi._ptr
    lad_T = None if ladder is None else np.ascontiguousarray(ladder, dtype=np.float64)
    lad_out = None if ladder is None else np.empty((n_sets, prob.n_meas, lad_T.size))
    rc = lib().trpl_emu_loglik_batch(prob.model, prob.n_meas, prob.meas, prob.n_times_total,
                                     p(prob.times, C.c_double), p(prob.vals, C.c_double),
                                     p(prob.uncs, C.c_double), p(prob.profiles, C.c_double), n_sets,
                                     p(params, C.c_double), p(aux, C.c_double), C.byref(opts),
                                     p(logll, C.c_double), p(status, C.c_int32), p(nsteps, C.c_int32),
                                     p(curves, C.c_double), p(prob.irf_moments, C.c_double),
                                     p(lad_T, C.c_double), 0 if ladder is None else lad_T.size, p(lad_out, C.c_double))
    if rc:
        raise RuntimeError("emu failed")
    if ladder is not None:
        return logll, status, nsteps, curves, lad_out
    return logll, status, nsteps, curves
