"""Builds tests/host_shim.cpp (the product's scalar headers compiled for the CPU) and wraps it with ctypes."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)
_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    out_dir = os.path.join(ROOT, "build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libpg_host_shim.so")
    srcs = [os.path.join(HERE, "host_shim.cpp"), os.path.join(ROOT, "pygemma_b200", "csrc", "pg_math.cuh"),
            os.path.join(ROOT, "pygemma_b200", "csrc", "pg_eval.cuh"),
            os.path.join(ROOT, "pygemma_b200", "csrc", "compress_plan.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", srcs[0], "-o", so, "-lm"])
    L = ctypes.CDLL(so)
    L.pgh_scan.restype = None
    L.pgh_scan.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_long, _dp, _dp, _dp, ctypes.c_int, ctypes.c_int,
                           _dp, _ip, _ip]
    L.pgh_f_sf.restype = ctypes.c_double
    L.pgh_f_sf.argtypes = [ctypes.c_double, ctypes.c_double]
    L.pgh_brent.restype = ctypes.c_double
    L.pgh_brent.argtypes = [_dp, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                            _dp, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    L.pgh_interp_error.restype = ctypes.c_double
    L.pgh_interp_error.argtypes = [ctypes.c_int, ctypes.c_int, _dp, _dp, ctypes.c_double, ctypes.c_int]
    L.pgh_plan_nodes.restype = ctypes.c_int
    L.pgh_plan_nodes.argtypes = [ctypes.c_int, _dp]
    L.pgh_fused_plan_check.restype = ctypes.c_double
    L.pgh_fused_plan_check.argtypes = [ctypes.c_int, _dp, _dp, ctypes.c_int, ctypes.c_int, _ip, _ip]
    L.pgh_compress_error.restype = ctypes.c_double
    L.pgh_compress_error.argtypes = [ctypes.c_int, _dp, _dp, ctypes.c_int]
    L.pgh_scan_compressed.restype = None
    L.pgh_scan_compressed.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_long, _dp, _dp, _dp, ctypes.c_int,
                                      ctypes.c_int, _dp, _ip, _ip, _ip]
    L.pgh_scan_compressed_ex.restype = None
    L.pgh_scan_compressed_ex.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_long, _dp, _dp, _dp, ctypes.c_int,
                                         ctypes.c_int, _dp, _ip, _ip, _ip, ctypes.c_int, _dp, _dp]
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(_dp)


def scan(d, y, w0, xr_snp_major, grid=False, exact_w0y=False):
    d = np.ascontiguousarray(d, dtype=np.float64)
    w0 = np.asarray(w0, dtype=np.float64)
    wy = np.asfortranarray(np.concatenate([w0, np.asarray(y, dtype=np.float64).reshape(-1, 1)], axis=1))
    xr = np.ascontiguousarray(xr_snp_major, dtype=np.float64)
    n, c0, m = d.shape[0], w0.shape[1], xr.shape[0]
    out = np.empty((m, 6))
    st = np.zeros(m, dtype=np.int32)
    ev = np.zeros((m, 2), dtype=np.int32)
    lib().pgh_scan(n, c0, m, _p(d), _p(wy), _p(xr), int(grid), int(exact_w0y), _p(out), st.ctypes.data_as(_ip),
                   ev.ctypes.data_as(_ip))
    cols = ["beta", "se_beta", "tau", "lambda", "F_wald", "p_wald"]
    r = {c: out[:, i].copy() for i, c in enumerate(cols)}
    r["status"] = st
    r["n_eval2"] = ev[:, 0].copy()
    r["n_eval3"] = ev[:, 1].copy()
    return r


def scan_compressed(d, y, w0, xr_snp_major, grid=False, xrow_form=True):
    """The scan evaluated from the eigenvalue-space compressed moments (compress_plan.h); xrow_form: covariate levels
    of the Pab recursion taken from the table-2 rows (pg_eval.cuh), as reml_solve_kernel does."""
    d = np.ascontiguousarray(d, dtype=np.float64)
    w0 = np.asarray(w0, dtype=np.float64)
    wy = np.asfortranarray(np.concatenate([w0, np.asarray(y, dtype=np.float64).reshape(-1, 1)], axis=1))
    xr = np.ascontiguousarray(xr_snp_major, dtype=np.float64)
    n, c0, m = d.shape[0], w0.shape[1], xr.shape[0]
    out = np.empty((m, 6))
    st = np.zeros(m, dtype=np.int32)
    ev = np.zeros((m, 2), dtype=np.int32)
    kc = np.zeros(1, dtype=np.int32)
    lib().pgh_scan_compressed(n, c0, m, _p(d), _p(wy), _p(xr), int(grid), int(xrow_form), _p(out), st.ctypes.data_as(_ip),
                              ev.ctypes.data_as(_ip), kc.ctypes.data_as(_ip))
    cols = ["beta", "se_beta", "tau", "lambda", "F_wald", "p_wald"]
    r = {c: out[:, i].copy() for i, c in enumerate(cols)}
    r["status"] = st
    r["n_eval2"] = ev[:, 0].copy()
    r["n_eval3"] = ev[:, 1].copy()
    r["nodes"] = int(kc[0])
    return r


def scan_compressed_ex(d, y, w0, xr_snp_major, swap=False, lrt=False):
    """x-row form of the compressed scan with the product's optional modes: swap = "de" role swap (PG_SCAN_DE),
    lrt = the ML optimiser of the likelihood-ratio outputs (pg::MlSolver) on the null and every alternative model."""
    d = np.ascontiguousarray(d, dtype=np.float64)
    w0 = np.asarray(w0, dtype=np.float64)
    wy = np.asfortranarray(np.concatenate([w0, np.asarray(y, dtype=np.float64).reshape(-1, 1)], axis=1))
    xr = np.ascontiguousarray(xr_snp_major, dtype=np.float64)
    n, c0, m = d.shape[0], w0.shape[1], xr.shape[0]
    out = np.empty((m, 6))
    st = np.zeros(m, dtype=np.int32)
    ev = np.zeros((m, 2), dtype=np.int32)
    kc = np.zeros(1, dtype=np.int32)
    null3 = np.empty(3)
    lrt2 = np.empty((m, 2))
    lib().pgh_scan_compressed_ex(n, c0, m, _p(d), _p(wy), _p(xr), 0, 1, _p(out), st.ctypes.data_as(_ip),
                                 ev.ctypes.data_as(_ip), kc.ctypes.data_as(_ip), int(swap), _p(null3) if lrt else None,
                                 _p(lrt2) if lrt else None)
    cols = ["beta", "se_beta", "tau", "lambda", "F_wald", "p_wald"]
    r = {c: out[:, i].copy() for i, c in enumerate(cols)}
    r["status"] = st
    if lrt:
        r.update(lambda_null=null3[0], tau_null=null3[1], l_null=null3[2], lambda_ml=lrt2[:, 0].copy(),
                 loglik_ml=lrt2[:, 1].copy())
    return r
