"""ctypes binding of libpygemma_b200.so (include/pygemma_b200.h).

The shared library is the product: if it is missing or does not load this
module raises -- there is no Python or CPU fallback for the scan.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PYGEMMA_B200_LIB selects another build of the same library (the bounds-checked debug build, csrc/pg_debug.cuh)
LIB_PATH = os.environ.get("PYGEMMA_B200_LIB") or os.path.join(_HERE, "libpygemma_b200.so")

PG_X_I8, PG_X_F32, PG_X_F64, PG_X_BED = 0, 1, 2, 3
PG_X_SAMPLE_MAJOR, PG_X_SNP_MAJOR = 0, 1
PG_ROT_AUTO, PG_ROT_FP64, PG_ROT_I8SPLIT, PG_ROT_I8TC = 0, 1, 2, 3
PG_ROT_I8TC_MOMENTS = 4   # reported by pg_timing.rot_engine only: the fused kernel also produced the moments
PG_REML_AUTO, PG_REML_COMPRESSED, PG_REML_STREAM, PG_REML_WARP = 0, 1, 2, 3
PG_SCAN_WALD, PG_SCAN_DE = 0, 1
PG_MAX_TRAITS = 64  # traits per pass of pg_set_design_multi (include/pygemma_b200.h)

# every symbol include/pygemma_b200.h declares (tests check the library exports each one)
SYMBOLS = [
    "pg_abi_version", "pg_rotation_planes", "pg_device_count", "pg_last_error", "pg_create", "pg_destroy", "pg_set_kinship",
    "pg_set_eigen", "pg_set_eigen_device", "pg_get_eigen_device", "pg_copy_eigen", "pg_set_design", "pg_set_design_multi", "pg_set_stream", "pg_set_options",
    "pg_set_reml_engine", "pg_set_scan_mode", "pg_grm", "pg_set_bed_options", "pg_set_moment_fusion", "pg_probe_fusion",
    "pg_probe_rotation_launches",
    "pg_scan", "pg_scan_device", "pg_scan_lrt", "pg_null_model",
    "pg_multi_create", "pg_multi_destroy", "pg_multi_last_error", "pg_multi_count", "pg_multi_handle", "pg_multi_set_kinship",
    "pg_multi_set_eigen", "pg_multi_set_design", "pg_multi_scan", "pg_probe_precompute", "pg_probe_f_sf", "pg_probe_rotated",
    "pg_probe_block_plan",
]


class PgTiming(ctypes.Structure):
    _fields_ = [("total_ms", ctypes.c_float), ("h2d_ms", ctypes.c_float), ("convert_ms", ctypes.c_float),
                ("rotate_ms", ctypes.c_float), ("reml_ms", ctypes.c_float), ("d2h_ms", ctypes.c_float),
                ("n_blocks", ctypes.c_int32), ("block_snps", ctypes.c_int32), ("reml_launches", ctypes.c_int32),
                ("rotate_launches", ctypes.c_int32), ("convert_launches", ctypes.c_int32),
                ("rot_engine", ctypes.c_int32), ("compress_ms", ctypes.c_float), ("reml_engine", ctypes.c_int32),
                ("n_nodes", ctypes.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ }


class PgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pygemma_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Load the CUDA library (building it is __graft_entry__.build()'s / pygemma_b200.build's job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m pygemma_b200.build` (nvcc, sm_100a). "
            "pygemma_b200 has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double
    L.pg_abi_version.restype = i32
    L.pg_rotation_planes.restype = i32
    L.pg_device_count.argtypes = [ctypes.POINTER(i32)]
    L.pg_last_error.restype = ctypes.c_char_p
    L.pg_last_error.argtypes = [vp]
    L.pg_create.argtypes = [i32, i32, i32, ctypes.POINTER(vp)]
    L.pg_destroy.argtypes = [vp]
    L.pg_set_kinship.argtypes = [vp, vp, vp, ctypes.POINTER(ctypes.c_float)]
    L.pg_set_eigen.argtypes = [vp, vp, i32, vp]
    L.pg_set_eigen_device.argtypes = [vp, vp, i32, vp]
    L.pg_get_eigen_device.argtypes = [vp, vp, vp]
    L.pg_copy_eigen.argtypes = [vp, vp]
    L.pg_set_design.argtypes = [vp, vp, vp, i32, ctypes.POINTER(ctypes.c_float)]
    L.pg_set_design_multi.argtypes = [vp, vp, vp, i32, i32, ctypes.POINTER(ctypes.c_float)]
    L.pg_set_options.argtypes = [vp, i32, i64]
    L.pg_set_stream.argtypes = [vp, vp]
    L.pg_set_reml_engine.argtypes = [vp, i32]
    L.pg_set_scan_mode.argtypes = [vp, i32]
    L.pg_set_moment_fusion.argtypes = [vp, i32]
    L.pg_probe_rotation_launches.argtypes = [i32, i64, i64, i64, i32, ctypes.POINTER(i32)]
    L.pg_probe_fusion.argtypes = [vp, ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(ctypes.c_float)]
    L.pg_set_bed_options.argtypes = [vp, i32, i32]
    L.pg_grm.argtypes = [vp, vp, i32, i64, i32, i64, vp, i32, vp, ctypes.POINTER(ctypes.c_float),
                         ctypes.POINTER(ctypes.c_float)]
    scan_args = [vp, vp, i32, i64, i32, i64, i32] + [vp] * 9 + [ctypes.POINTER(PgTiming)]
    L.pg_scan.argtypes = scan_args
    L.pg_scan_device.argtypes = scan_args
    L.pg_scan_lrt.argtypes = [vp, vp, i32, i64, i32, i64, i32] + [vp] * 13 + [ctypes.POINTER(PgTiming)]
    L.pg_null_model.argtypes = [vp, i32, vp, vp, vp]
    L.pg_probe_precompute.argtypes = [vp, vp, dbl, i32, i32, vp]
    L.pg_probe_f_sf.argtypes = [vp, vp, dbl, i64, vp]
    L.pg_probe_rotated.argtypes = [vp, vp, i64, ctypes.POINTER(i64)]
    L.pg_probe_block_plan.argtypes = [i64, i64, i32, i32, ctypes.POINTER(i64), i32]
    fp = ctypes.POINTER(ctypes.c_float)
    L.pg_multi_create.argtypes = [i32, i32, i32, ctypes.POINTER(i32), ctypes.POINTER(vp)]
    L.pg_multi_destroy.argtypes = [vp]
    L.pg_multi_last_error.argtypes = [vp]
    L.pg_multi_count.argtypes = [vp]
    L.pg_multi_handle.argtypes = [vp, i32]
    L.pg_multi_set_kinship.argtypes = [vp, vp, vp, fp, fp]
    L.pg_multi_set_eigen.argtypes = [vp, vp, i32, vp]
    L.pg_multi_set_design.argtypes = [vp, vp, vp, i32, fp]
    L.pg_multi_scan.argtypes = [vp, vp, i32, i64, i32, i64, i32] + [vp] * 9 + [ctypes.POINTER(PgTiming)]
    for name in SYMBOLS:
        if name not in ("pg_last_error", "pg_multi_last_error", "pg_multi_handle"):
            getattr(L, name).restype = i32
    L.pg_last_error.restype = ctypes.c_char_p
    L.pg_multi_last_error.restype = ctypes.c_char_p
    L.pg_multi_handle.restype = vp
    _lib = L
    return L


def block_plan(m: int, blk: int, host_input: bool = True, fixed_block: bool = False):
    """SNP-block boundaries pg_scan uses (pg_probe_block_plan; host arithmetic, no device needed)."""
    L = load()
    cap = 4096
    buf = (ctypes.c_int64 * cap)()
    nb = L.pg_probe_block_plan(int(m), int(blk), int(bool(host_input)), int(bool(fixed_block)), buf, cap)
    if nb < 0:
        raise PgError(nb, "pg_probe_block_plan: bad arguments")
    return [int(buf[i]) for i in range(min(nb + 1, cap))] if m else []


def _ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


_XDT = {np.dtype(np.int8): PG_X_I8, np.dtype(np.float32): PG_X_F32, np.dtype(np.float64): PG_X_F64}


def xdtype_of(a) -> int:
    try:
        return _XDT[np.dtype(a.dtype)]
    except KeyError:
        raise TypeError(f"genotype dtype {a.dtype} not supported by the C ABI (int8, float32, float64)")


class Handle:
    """Owns one pg_handle (one GPU).  Thin, 1:1 with the C ABI."""

    def __init__(self, n: int, c0: int, device: int = 0):
        self.L = load()
        self.n, self.c0, self.device = int(n), int(c0), int(device)
        self.q = 1   # phenotypes of the current design (set_design with an (n, q) matrix)
        h = ctypes.c_void_p()
        rc = self.L.pg_create(self.n, self.c0, self.device, ctypes.byref(h))
        if rc != 0:
            raise PgError(rc, self.L.pg_last_error(None).decode())
        self.h = h

    def _ck(self, rc):
        if rc != 0:
            raise PgError(rc, self.L.pg_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.pg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # --- setup -------------------------------------------------------------------------------
    def set_kinship(self, K):
        K = np.ascontiguousarray(K, dtype=np.float64)
        assert K.shape == (self.n, self.n)
        d = np.empty(self.n)
        ms = ctypes.c_float(0)
        self._ck(self.L.pg_set_kinship(self.h, _ptr(K), _ptr(d), ctypes.byref(ms)))
        return d, float(ms.value)

    def grm(self, X, layout=PG_X_SAMPLE_MAJOR, return_K=True, set_kinship=False):
        """Genetic relatedness matrix K = Z Z^T / p of the marker matrix X on the device (pg_grm); optionally
        eigendecomposed in place for the scan.  Returns dict(K, d, grm_ms, eig_ms)."""
        if X.ndim != 2:
            raise ValueError("X must be 2-D")
        n, p = X.shape if layout == PG_X_SAMPLE_MAJOR else X.shape[::-1]
        if n != self.n:
            raise ValueError(f"X has {n} samples, handle was created for {self.n}")
        if X.strides[1] != X.itemsize or (X.shape[0] > 1 and X.strides[0] % X.itemsize) or X.strides[0] < 0:
            X = np.ascontiguousarray(X)
        ld = max(X.strides[0] // X.itemsize if X.shape[0] > 1 else X.shape[1], X.shape[1])
        K = np.empty((self.n, self.n)) if return_K else None
        d = np.empty(self.n) if set_kinship else None
        gms, ems = ctypes.c_float(0), ctypes.c_float(0)
        self._ck(self.L.pg_grm(self.h, _ptr(X), xdtype_of(X), ld, layout, p, _ptr(K), int(bool(set_kinship)), _ptr(d),
                               ctypes.byref(gms), ctypes.byref(ems)))
        return {"K": K, "d": d, "grm_ms": float(gms.value), "eig_ms": float(ems.value)}

    def set_eigen(self, U, d):
        d = np.ascontiguousarray(d, dtype=np.float64).reshape(-1)
        assert d.shape[0] == self.n
        if U is None:
            self._ck(self.L.pg_set_eigen(self.h, None, 0, _ptr(d)))
            return
        U = np.asarray(U, dtype=np.float64)
        assert U.shape == (self.n, self.n)
        if U.flags.f_contiguous and not U.flags.c_contiguous:
            self._ck(self.L.pg_set_eigen(self.h, _ptr(U), 0, _ptr(d)))
        else:
            U = np.ascontiguousarray(U)
            self._ck(self.L.pg_set_eigen(self.h, _ptr(U), 1, _ptr(d)))

    def set_eigen_device(self, U_ptr: int, u_row_major: bool, d_ptr: int):
        self._ck(self.L.pg_set_eigen_device(self.h, ctypes.c_void_p(U_ptr), int(u_row_major), ctypes.c_void_p(d_ptr)))

    def get_eigen_device(self, U_ptr: int, d_ptr: int):
        self._ck(self.L.pg_get_eigen_device(self.h, ctypes.c_void_p(U_ptr), ctypes.c_void_p(d_ptr)))

    def copy_eigen_from(self, other: "Handle"):
        """Take over `other`'s eigen-system (U, d) device to device (pg_copy_eigen)."""
        self._ck(self.L.pg_copy_eigen(self.h, other.h))

    def set_design(self, W, y, already_rotated=False):
        """y: n values, or an (n, q) matrix of q phenotypes scanned together (outputs become (q, m) arrays)."""
        W = np.ascontiguousarray(W, dtype=np.float64).reshape(self.n, self.c0)
        y = np.ascontiguousarray(y, dtype=np.float64)
        ms = ctypes.c_float(0)
        if y.ndim == 2 and y.shape[1] > 1:
            assert y.shape[0] == self.n
            self._ck(self.L.pg_set_design_multi(self.h, _ptr(W) if self.c0 else None, _ptr(y), int(y.shape[1]),
                                                int(already_rotated), ctypes.byref(ms)))
            self.q = int(y.shape[1])
            return float(ms.value)
        y = y.reshape(-1)
        assert y.shape[0] == self.n
        self._ck(self.L.pg_set_design(self.h, _ptr(W) if self.c0 else None, _ptr(y), int(already_rotated),
                                      ctypes.byref(ms)))
        self.q = 1
        return float(ms.value)

    def set_stream(self, stream_ptr: int):
        """Run the handle's device work on a caller stream (e.g. torch.cuda.current_stream().cuda_stream)."""
        self._ck(self.L.pg_set_stream(self.h, ctypes.c_void_p(stream_ptr or None)))

    def set_options(self, rotation=PG_ROT_AUTO, block_snps=0):
        self._ck(self.L.pg_set_options(self.h, int(rotation), int(block_snps)))

    def set_reml_engine(self, engine=PG_REML_AUTO):
        self._ck(self.L.pg_set_reml_engine(self.h, int(engine)))

    def set_moment_fusion(self, mode=-1):
        """-1: the fused rotation also produces the moments where that pays; 0: never; 1: wherever the engine allows."""
        self._ck(self.L.pg_set_moment_fusion(self.h, int(mode)))

    def fusion_info(self):
        """What the next scan of the current design would do: {'fused', 'g_columns', 'pieces', 'build_ms'}."""
        a, b, c = ctypes.c_int32(0), ctypes.c_int32(0), ctypes.c_int32(0)
        ms = ctypes.c_float(0)
        self._ck(self.L.pg_probe_fusion(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(ms)))
        return {"fused": bool(a.value), "g_columns": b.value, "pieces": c.value, "build_ms": ms.value}

    def set_scan_mode(self, mode=PG_SCAN_WALD):
        """PG_SCAN_DE: scanned columns are phenotypes, the design's y is the tested regressor (reference de=True)."""
        self._ck(self.L.pg_set_scan_mode(self.h, int(mode)))

    # --- scan --------------------------------------------------------------------------------
    def null_model(self, trait=0):
        """ML fit of the null model [W] of one phenotype of the current design: dict(lambda_null, tau_null, l_null)."""
        v = (ctypes.c_double * 3)()
        self._ck(self.L.pg_null_model(self.h, int(trait), ctypes.byref(v, 0), ctypes.byref(v, 8), ctypes.byref(v, 16)))
        return {"lambda_null": v[0], "tau_null": v[1], "l_null": v[2]}

    def scan(self, X, grid=False, layout=PG_X_SAMPLE_MAJOR, with_counts=True, lrt=False):
        """Host-buffer scan.  X: (n, m) sample-major or (m, n) SNP-major ndarray, any stride along rows.
        lrt: also return lambda_ml, loglik_ml, D_lrt, p_lrt (pg_scan_lrt)."""
        if X.ndim != 2:
            raise ValueError("X must be 2-D")
        if layout == PG_X_SAMPLE_MAJOR:
            n, m = X.shape
        else:
            m, n = X.shape
        if n != self.n:
            raise ValueError(f"X has {n} samples, handle was created for {self.n}")
        if X.strides[1] != X.itemsize or (X.shape[0] > 1 and X.strides[0] % X.itemsize) or X.strides[0] < 0:
            X = np.ascontiguousarray(X)
        ld = X.strides[0] // X.itemsize if X.shape[0] > 1 else X.shape[1]
        ld = max(ld, X.shape[1])
        shape = m if self.q == 1 else (self.q, m)
        out = {k: np.empty(shape) for k in ("beta", "se_beta", "tau", "lambda", "F_wald", "p_wald")}
        st = np.zeros(shape, dtype=np.int32)
        e2 = np.zeros(shape, dtype=np.int32) if with_counts else None
        e3 = np.zeros(shape, dtype=np.int32) if with_counts else None
        tm = PgTiming()
        if lrt:
            out.update({k: np.empty(shape) for k in ("lambda_ml", "loglik_ml", "D_lrt", "p_lrt")})
            self._ck(self.L.pg_scan_lrt(self.h, _ptr(X), xdtype_of(X), ld, layout, m, int(bool(grid)),
                                        _ptr(out["beta"]), _ptr(out["se_beta"]), _ptr(out["tau"]), _ptr(out["lambda"]),
                                        _ptr(out["F_wald"]), _ptr(out["p_wald"]), _ptr(out["lambda_ml"]),
                                        _ptr(out["loglik_ml"]), _ptr(out["D_lrt"]), _ptr(out["p_lrt"]), _ptr(st), _ptr(e2),
                                        _ptr(e3), ctypes.byref(tm)))
        else:
            self._ck(self.L.pg_scan(self.h, _ptr(X), xdtype_of(X), ld, layout, m, int(bool(grid)),
                                    _ptr(out["beta"]), _ptr(out["se_beta"]), _ptr(out["tau"]), _ptr(out["lambda"]),
                                    _ptr(out["F_wald"]), _ptr(out["p_wald"]), _ptr(st), _ptr(e2), _ptr(e3),
                                    ctypes.byref(tm)))
        out["status"] = st
        if with_counts:
            out["n_eval2"], out["n_eval3"] = e2, e3
        out["timing"] = tm.as_dict()
        return out

    def scan_bed(self, packed, grid=False, count_A1=False, standardize=False, with_counts=True):
        """Scan a packed PLINK .bed body: `packed` is a C-contiguous (m, bytes_per_snp) uint8 array, bytes_per_snp >=
        ceil(n/4).  Missing genotypes are mean-imputed (and the columns optionally standardised) on the device."""
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        if packed.ndim != 2 or packed.shape[1] < (self.n + 3) // 4:
            raise ValueError(f"packed .bed block must have shape (m, >= {(self.n + 3) // 4})")
        m = packed.shape[0]
        self._ck(self.L.pg_set_bed_options(self.h, int(bool(count_A1)), int(bool(standardize))))
        shape = m if self.q == 1 else (self.q, m)
        out = {k: np.empty(shape) for k in ("beta", "se_beta", "tau", "lambda", "F_wald", "p_wald")}
        st = np.zeros(shape, dtype=np.int32)
        e2 = np.zeros(shape, dtype=np.int32) if with_counts else None
        e3 = np.zeros(shape, dtype=np.int32) if with_counts else None
        tm = PgTiming()
        self._ck(self.L.pg_scan(self.h, _ptr(packed), PG_X_BED, packed.shape[1], PG_X_SNP_MAJOR, m, int(bool(grid)),
                                _ptr(out["beta"]), _ptr(out["se_beta"]), _ptr(out["tau"]), _ptr(out["lambda"]),
                                _ptr(out["F_wald"]), _ptr(out["p_wald"]), _ptr(st), _ptr(e2), _ptr(e3),
                                ctypes.byref(tm)))
        out["status"] = st
        if with_counts:
            out["n_eval2"], out["n_eval3"] = e2, e3
        out["timing"] = tm.as_dict()
        return out

    def scan_device(self, x_ptr: int, xdtype: int, ld: int, layout: int, m: int, grid: bool, out_ptrs, status_ptr=0,
                    e2_ptr=0, e3_ptr=0):
        """Device-buffer scan: x_ptr and the six out_ptrs are raw CUDA device pointers."""
        tm = PgTiming()
        vp = ctypes.c_void_p
        self._ck(self.L.pg_scan_device(self.h, vp(x_ptr), xdtype, ld, layout, m, int(bool(grid)),
                                       *[vp(p) for p in out_ptrs], vp(status_ptr or None), vp(e2_ptr or None),
                                       vp(e3_ptr or None), ctypes.byref(tm)))
        return tm.as_dict()

    # --- probes ------------------------------------------------------------------------------
    def probe_precompute(self, x_rot, lam, fixed_index=-1, full=True):
        x = np.ascontiguousarray(x_rot, dtype=np.float64).reshape(-1)
        out = np.empty(9)
        self._ck(self.L.pg_probe_precompute(self.h, _ptr(x), float(lam), int(fixed_index), int(bool(full)), _ptr(out)))
        return out

    def probe_f_sf(self, F, nu):
        F = np.ascontiguousarray(F, dtype=np.float64).reshape(-1)
        p = np.empty_like(F)
        self._ck(self.L.pg_probe_f_sf(self.h, _ptr(F), float(nu), F.shape[0], _ptr(p)))
        return p

    def probe_rotated(self, count):
        xr = np.empty((count, self.n))
        row0 = ctypes.c_int64(0)
        self._ck(self.L.pg_probe_rotated(self.h, _ptr(xr), count, ctypes.byref(row0)))
        return xr, int(row0.value)


class MultiHandle:
    """Several GPUs of one node driven by this process (pg_multi_*): one pg_handle per device, U and d copied peer to peer,
    contiguous SNP shards scanned concurrently, rows written in input order.  One phenotype per pass."""

    def __init__(self, n: int, c0: int, devices):
        self.L = load()
        devices = list(range(int(devices))) if isinstance(devices, (int, np.integer)) else [int(d) for d in devices]
        if not devices:
            raise ValueError("MultiHandle needs at least one device")
        self.n, self.c0, self.devices, self.q = int(n), int(c0), devices, 1
        arr = (ctypes.c_int * len(devices))(*devices)
        h = ctypes.c_void_p()
        rc = self.L.pg_multi_create(self.n, self.c0, len(devices), arr, ctypes.byref(h))
        if rc != 0:
            raise PgError(rc, self.L.pg_multi_last_error(None).decode())
        self.h = h

    def _ck(self, rc):
        if rc != 0:
            raise PgError(rc, self.L.pg_multi_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.pg_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_kinship(self, K):
        K = np.ascontiguousarray(K, dtype=np.float64)
        assert K.shape == (self.n, self.n)
        d = np.empty(self.n)
        ms, bms = ctypes.c_float(0), ctypes.c_float(0)
        self._ck(self.L.pg_multi_set_kinship(self.h, _ptr(K), _ptr(d), ctypes.byref(ms), ctypes.byref(bms)))
        self.bcast_ms = float(bms.value)
        return d, float(ms.value)

    def set_eigen(self, U, d):
        d = np.ascontiguousarray(d, dtype=np.float64).reshape(-1)
        if U is None:
            self._ck(self.L.pg_multi_set_eigen(self.h, None, 0, _ptr(d)))
            return
        U = np.asarray(U, dtype=np.float64)
        if U.flags.f_contiguous and not U.flags.c_contiguous:
            self._ck(self.L.pg_multi_set_eigen(self.h, _ptr(U), 0, _ptr(d)))
        else:
            U = np.ascontiguousarray(U)
            self._ck(self.L.pg_multi_set_eigen(self.h, _ptr(U), 1, _ptr(d)))

    def set_design(self, W, y, already_rotated=False):
        W = np.ascontiguousarray(W, dtype=np.float64).reshape(self.n, self.c0)
        y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
        assert y.shape[0] == self.n
        ms = ctypes.c_float(0)
        self._ck(self.L.pg_multi_set_design(self.h, _ptr(W) if self.c0 else None, _ptr(y), int(already_rotated), ctypes.byref(ms)))
        return float(ms.value)

    def set_moment_fusion(self, mode=-1):
        """-1: the fused rotation also produces the moments where that pays; 0: never; 1: wherever the engine allows."""
        self._ck(self.L.pg_set_moment_fusion(self.h, int(mode)))

    def fusion_info(self):
        """What the next scan of the current design would do: {'fused', 'g_columns', 'pieces', 'build_ms'}."""
        a, b, c = ctypes.c_int32(0), ctypes.c_int32(0), ctypes.c_int32(0)
        ms = ctypes.c_float(0)
        self._ck(self.L.pg_probe_fusion(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(ms)))
        return {"fused": bool(a.value), "g_columns": b.value, "pieces": c.value, "build_ms": ms.value}

    def set_scan_mode(self, mode=PG_SCAN_WALD):
        for i in range(len(self.devices)):
            rc = self.L.pg_set_scan_mode(ctypes.c_void_p(self.L.pg_multi_handle(self.h, i)), int(mode))
            if rc != 0:
                raise PgError(rc, "pg_set_scan_mode")

    def scan(self, X, grid=False, layout=PG_X_SAMPLE_MAJOR, with_counts=True, lrt=False):
        if lrt:
            raise ValueError("the likelihood-ratio outputs are not available through MultiHandle")
        if X.ndim != 2:
            raise ValueError("X must be 2-D")
        n, m = X.shape if layout == PG_X_SAMPLE_MAJOR else X.shape[::-1]
        if n != self.n:
            raise ValueError(f"X has {n} samples, handle was created for {self.n}")
        if X.strides[1] != X.itemsize or (X.shape[0] > 1 and X.strides[0] % X.itemsize) or X.strides[0] < 0:
            X = np.ascontiguousarray(X)
        ld = max(X.strides[0] // X.itemsize if X.shape[0] > 1 else X.shape[1], X.shape[1])
        out = {k: np.empty(m) for k in ("beta", "se_beta", "tau", "lambda", "F_wald", "p_wald")}
        st = np.zeros(m, dtype=np.int32)
        e2 = np.zeros(m, dtype=np.int32) if with_counts else None
        e3 = np.zeros(m, dtype=np.int32) if with_counts else None
        tms = (PgTiming * len(self.devices))()
        self._ck(self.L.pg_multi_scan(self.h, _ptr(X), xdtype_of(X), ld, layout, m, int(bool(grid)),
                                      _ptr(out["beta"]), _ptr(out["se_beta"]), _ptr(out["tau"]), _ptr(out["lambda"]),
                                      _ptr(out["F_wald"]), _ptr(out["p_wald"]), _ptr(st), _ptr(e2), _ptr(e3), tms))
        out["status"] = st
        if with_counts:
            out["n_eval2"], out["n_eval3"] = e2, e3
        per = [t.as_dict() for t in tms]
        out["timing_per_device"] = per
        out["timing"] = max(per, key=lambda t: t["total_ms"])   # the slowest device is the step
        return out


def device_count() -> int:
    c = ctypes.c_int(0)
    load().pg_device_count(ctypes.byref(c))
    return int(c.value)
