"""Drop-in for the reference's `pygemma.lmm.pygemma` (reference lmm/lmm.py:87-411), B200-native.

    from pygemma_b200 import lmm
    df = lmm.pygemma(Y, X, W, K, snps=snps, verbose=1)

Same call, same argument meaning, same DataFrame (`beta, se_beta, tau, lambda, F_wald, p_wald[, SNPs]`,
one row per genotype column in input order, nothing filtered).  The work happens in
libpygemma_b200.so (CUDA, sm_100a) through the C ABI of include/pygemma_b200.h:

    eigh(K)                 lmm/lmm.py:151-162  -> cuSOLVER syevd on the GPU (timed separately)
    U.T @ {X, Y, W}         lmm/lmm.py:243-246  -> device GEMMs, X in SNP blocks
    Pool.imap(calculate)    lmm/lmm.py:378-403  -> per-SNP REML kernel

Extensions (all opt-in): `F = lmm.factorize(K)` decomposes once and `lmm.pygemma(Y, X, W, F)` reuses it (KinshipFactor);
`lmm.pygemma_multi` scans several traits per pass; `lrt=True` adds the likelihood-ratio columns of the reference's
commented-out scaffolding; `gpus=N` drives several GPUs from this one process (under torchrun every rank owns one GPU and
the call shards by itself); `lmm.null_model` returns the null-model fit.

Differences from the reference, all documented in DESIGN.md: arithmetic is float64 throughout (the
reference mixes float32 storage into a float64 core), so every returned column is float64; `nproc` is
accepted and ignored (parallelism is the GPU, or several GPUs under torch.distributed); `de=True` runs the
role-swapped scan that calculate_de spells out (lmm/lmm.py:498-532: each column of X is the outcome, Y the predictor) --
the reference's own driver raises before reaching it (lmm/lmm.py:499 unpacks SampleIter's 5-tuple into 4 names).
There is no CPU fallback: without the CUDA library or a GPU the call raises.
"""
from __future__ import annotations

import time

import numpy as np
import pandas as pd

from . import _capi

COLUMNS = ["beta", "se_beta", "tau", "lambda", "F_wald", "p_wald"]
# opt-in likelihood-ratio columns, named as the reference's commented-out result keys (lmm/lmm.py:137-141)
LRT_COLUMNS = ["D_lrt", "p_lrt", "likelihood"]

# null model of the most recent lrt=True call: one dict(lambda_null, tau_null, l_null) per trait (lmm/lmm.py:176-190)
last_null_model: list = []

# timings of the most recent call (seconds / milliseconds), for callers that want them without verbose
last_timing: dict = {}


def _as_genotypes(X):
    X = np.asarray(X)
    if X.ndim != 2:
        raise ValueError("X must have shape (n, m)")
    dt = X.dtype
    if dt == np.int8 or dt == np.float32 or dt == np.float64:
        return X
    if dt == np.bool_:
        return X.astype(np.int8)
    if np.issubdtype(dt, np.integer):
        if X.size == 0 or (X.min() >= -128 and X.max() <= 127):
            return X.astype(np.int8)
        return X.astype(np.float64)
    if dt == np.float16:
        return X.astype(np.float32)
    return X.astype(np.float64)


class KinshipFactor:
    """The eigendecomposition K = U diag(d) U^T, done once and kept in HBM (extension).

    The reference runs `eigh(K)` inside every `lmm.pygemma` call (lmm/lmm.py:151-162) and its callers loop that call
    over phenotypes and covariate sets (experiments/animal_gwas/run_gwas.py:167-175, benchmark_pygemma.py:238).  Pass
    the object returned by `factorize(K)` in the K position instead of the matrix and every call skips the
    decomposition: U, its int8 digit planes and the eigenvalue-space compression plan stay on the device.

        F = lmm.factorize(K)
        for y in traits: df = lmm.pygemma(y, X, W, F)

    Under torch.distributed (backend nccl) rank 0 decomposes and U, d are broadcast once; K may be None elsewhere.
    """

    def __init__(self, K, Z=None, device=None, c0=1, verbose=0):
        from . import multi

        self.ctx = multi.context(device)
        multi.check_backend(self.ctx)
        if K is not None:
            K = np.asarray(K, dtype=np.float64)
            if Z is not None:
                Z = np.asarray(Z, dtype=np.float64)
                K = Z @ K @ Z.T  # lmm/lmm.py:124-125
            if K.ndim != 2 or K.shape[0] != K.shape[1]:
                raise ValueError(f"K must be square, got {K.shape}")
        elif self.ctx.rank == 0:
            raise ValueError("K is required on rank 0")
        n = multi.agree_on_n(self.ctx, None if K is None else K.shape[0])
        self.n = int(n)
        self._h = _capi.Handle(self.n, int(c0), self.ctx.device)
        t0 = time.time()
        _log(verbose, "Starting eigendecomposition...")
        self.eig_ms = multi.setup_eigen(self.ctx, self._h, K)
        self.setup_s = time.time() - t0
        _log(verbose, f"Eigendecomposition computed - {round(self.setup_s, 3)} s")

    def handle(self, c0: int):
        """The live handle for designs with c0 covariate columns (re-created device to device when c0 changes)."""
        if self._h is None:
            raise ValueError("this KinshipFactor was closed")
        if self._h.c0 != int(c0):
            new = _capi.Handle(self.n, int(c0), self.ctx.device)
            try:
                new.copy_eigen_from(self._h)
            except Exception:
                new.close()
                raise
            self._h.close()
            self._h = new
        return self._h

    def close(self):
        if getattr(self, "_h", None) is not None:
            self._h.close()
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def factorize(K, Z=None, device=None, c0=1, verbose=0) -> KinshipFactor:
    """eigh(K) once, on the GPU, for any number of later `pygemma(Y, X, W, factor)` calls (see KinshipFactor).
    `c0` is a hint (the covariate count of the first call) that saves one handle re-creation."""
    return KinshipFactor(K, Z=Z, device=device, c0=c0, verbose=verbose)


def _log(verbose, msg):
    if verbose > 0:
        print(f"[pygemma_b200] {msg}", flush=True)


def pygemma(Y, X, W, K, Z=None, snps=None, verbose=0, disable_checks=True, de=False, grid=False, eigen=True,
            nproc=1, device=None, lrt=False, gpus=None):
    """Per-SNP LMM association scan (GEMMA Wald test) on a B200.

    Args mirror the reference (lmm/lmm.py:87-103): Y (n,) or (n, 1) phenotype; X (n, m) genotypes
    (int8 dosages, float32 or float64; other dtypes are converted); W (n, c) covariates, intercept
    column supplied by the caller; K (n, n) relatedness matrix, or the (n,) eigenvalue vector when
    eigen=False (then Y, X, W are taken as already rotated, lmm/lmm.py:164-167); Z: K <- Z K Z^T;
    snps: labels for the 'SNPs' column; grid: 12-point grid search for lambda instead of
    Brent + Newton; de: differential-expression mode (X columns are outcomes, Y the predictor; grid is not
    forwarded, as in lmm/lmm.py:504).  `device` (extension) selects the CUDA device; default 0 or the
    torch.distributed local rank.  `lrt=True` (extension) adds the likelihood-ratio columns the reference keeps as
    commented-out scaffolding (lmm/lmm.py:137-141,:176-190,:278-300): 'D_lrt' = 2 (l_alt - l_null) with the ML fits of
    lmm.calc_lambda, 'p_lrt' = its chi-square(1) tail, 'likelihood' = l_alt; the null fit is left in `last_null_model`.
    `gpus` (extension): an int or a list of CUDA devices to be driven by THIS process (pg_multi_*: U and d copied peer to
    peer, contiguous SNP shards scanned concurrently, same rows in the same order); the one-process-per-GPU route under
    torchrun needs no argument.
    """
    Y = np.asarray(Y, dtype=np.float64).reshape(-1, 1)  # lmm/lmm.py:115-116
    if gpus is not None and not (isinstance(gpus, (int, np.integer)) and int(gpus) <= 1):
        return _run_multi_device(Y, X, W, K, Z, snps, verbose, disable_checks, de, grid, eigen, gpus, lrt)
    return _run(Y, X, W, K, Z, snps, verbose, disable_checks, de, grid, eigen, device, lrt)[0]


def _run_multi_device(Y, X, W, K, Z, snps, verbose, disable_checks, de, grid, eigen, gpus, lrt):
    """lmm.pygemma on several GPUs of one node from one process (include/pygemma_b200.h: pg_multi_*)."""
    global last_timing
    t_start = time.time()
    if lrt or isinstance(K, KinshipFactor):
        raise ValueError("gpus=... takes a kinship matrix (or eigenvalues with eigen=False) and has no lrt=True")
    if _dist_active():
        raise ValueError("gpus=... drives several devices from one process; under torch.distributed every rank owns one GPU")
    X = _as_genotypes(X)
    W = np.asarray(W, dtype=np.float64)
    if W.ndim == 1:
        W = W.reshape(-1, 1)
    n, m = X.shape
    c0 = W.shape[1]
    if Y.shape[0] != n or W.shape[0] != n:
        raise ValueError(f"shape mismatch: X {X.shape}, Y {Y.shape}, W {W.shape}")
    K = np.asarray(K, dtype=np.float64)
    if Z is not None:
        Zm = np.asarray(Z, dtype=np.float64)
        K = Zm @ K @ Zm.T  # lmm/lmm.py:124-125
    if eigen and K.shape != (n, n):
        raise ValueError(f"K must be ({n}, {n}) when eigen=True, got {K.shape}")
    if not disable_checks:
        if (X.dtype.kind == "f" and np.isnan(X).any()) or np.isnan(Y).any() or np.isnan(W).any():
            raise ValueError("NaNs present in data")
    if de:
        grid = False
    timing = {"n": n, "m": m, "c0": c0, "q": 1}
    with _capi.MultiHandle(n, c0, gpus) as mh:
        timing["devices"] = mh.devices
        if eigen:
            _log(verbose, "Starting eigendecomposition...")
            _, timing["eig_ms"] = mh.set_kinship(K)
            timing["bcast_ms"] = mh.bcast_ms
            _log(verbose, f"Eigendecomposition computed, U and d on {len(mh.devices)} devices - {round(time.time() - t_start, 3)} s")
        else:
            mh.set_eigen(None, K.reshape(-1))
        mh.set_scan_mode(_capi.PG_SCAN_DE if de else _capi.PG_SCAN_WALD)
        timing["design_ms"] = mh.set_design(W, Y.reshape(-1), already_rotated=not eigen)
        _log(verbose, f"Running {m} SNPs with {n} individuals on {len(mh.devices)} GPUs...")
        t0 = time.time()
        out = mh.scan(X, grid=grid)
        timing["scan_wall_s"] = time.time() - t0
        timing["scan"] = out["timing"]
    bad = (out["status"] & 1) != 0
    data = {}
    for c in COLUMNS:
        col = out[c]
        if bad.any():
            col = col.copy()
            col[bad] = np.nan
        data[c] = col
    df = pd.DataFrame(data, columns=COLUMNS)
    if snps is not None:
        df["SNPs"] = snps  # lmm/lmm.py:408-409
    timing["total_s"] = time.time() - t_start
    last_timing = timing
    return df


def _dist_active() -> bool:
    from . import multi

    return multi._dist() is not None


def null_model(Y, W, K, Z=None, eigen=True, device=None):
    """The null model of the reference's likelihood-ratio scaffolding (lmm/lmm.py:176-190): ML lambda as lmm.calc_lambda
    finds it for y ~ W, tau_null = n / y^T P y and the log-likelihood.  K may be a KinshipFactor.  Returns a dict."""
    from . import multi

    Y = np.asarray(Y, dtype=np.float64).reshape(-1)
    W = np.asarray(W, dtype=np.float64)
    if W.ndim == 1:
        W = W.reshape(-1, 1)
    n, c0 = W.shape
    if Y.shape[0] != n:
        raise ValueError(f"shape mismatch: Y {Y.shape}, W {W.shape}")
    if isinstance(K, KinshipFactor):
        if not eigen or Z is not None:
            raise ValueError("a KinshipFactor already holds U (apply Z in lmm.factorize; eigen=False does not apply)")
        h = K.handle(c0)
        h.set_design(W, Y)
        return h.null_model(0)
    ctx = multi.context(device)
    with _capi.Handle(n, c0, ctx.device) as h:
        if eigen:
            multi.check_backend(ctx)
            Km = None
            if ctx.rank == 0:
                Km = np.asarray(K, dtype=np.float64)
                if Z is not None:
                    Zm = np.asarray(Z, dtype=np.float64)
                    Km = Zm @ Km @ Zm.T
            multi.setup_eigen(ctx, h, Km)
        else:
            h.set_eigen(None, np.asarray(K, dtype=np.float64).reshape(-1))
        h.set_design(W, Y, already_rotated=not eigen)
        return h.null_model(0)


def pygemma_multi(Y, X, W, K, Z=None, snps=None, verbose=0, disable_checks=True, grid=False, eigen=True, device=None,
                  lrt=False):
    """Several phenotypes in one pass (extension; the reference is called once per trait, e.g.
    experiments/benchmarks/benchmarks.py loops lmm.pygemma over simulated phenotypes).

    Y is (n, q); every other argument is as in `pygemma`.  Returns a list of q DataFrames, the i-th identical (bit for
    bit) to `pygemma(Y[:, i], X, W, K, ...)`: eigh(K) and the rotation U^T X, which dominate a single-trait scan, are
    done once, and only the REML stage runs per trait.
    """
    Y = np.asarray(Y, dtype=np.float64)
    if Y.ndim == 1:
        Y = Y.reshape(-1, 1)
    if Y.ndim != 2:
        raise ValueError("Y must be (n, q)")
    return _run(Y, X, W, K, Z, snps, verbose, disable_checks, False, grid, eigen, device, lrt)


def _run(Y, X, W, K, Z, snps, verbose, disable_checks, de, grid, eigen, device, lrt=False):
    global last_timing, last_null_model
    t_start = time.time()
    if de:
        # calculate_de (lmm/lmm.py:498-532): every column of X is a phenotype, Y the tested regressor; it calls
        # calc_lambda_restricted without `grid` (:504), i.e. always Brent + Newton
        if Y.shape[1] != 1:
            raise ValueError("de=True takes one predictor Y")
        if lrt:
            raise ValueError("lrt=True is not available with de=True")
        grid = False

    from . import multi  # torch.distributed plumbing, only active when a process group exists

    X = _as_genotypes(X)
    q = Y.shape[1]
    W = np.asarray(W, dtype=np.float64)
    if W.ndim == 1:
        W = W.reshape(-1, 1)
    n, m = X.shape
    c0 = W.shape[1]
    if Y.shape[0] != n or W.shape[0] != n:
        raise ValueError(f"shape mismatch: X {X.shape}, Y {Y.shape}, W {W.shape}")
    factor = K if isinstance(K, KinshipFactor) else None
    ctx = factor.ctx if factor is not None else multi.context(device)
    if factor is not None:
        if not eigen:
            raise ValueError("a KinshipFactor holds U: it cannot be combined with eigen=False (pre-rotated inputs)")
        if Z is not None:
            raise ValueError("apply Z when factorizing: lmm.factorize(K, Z=Z)")
        if factor.n != n:
            raise ValueError(f"the KinshipFactor was built for n={factor.n}, X has {n} samples")
    else:
        if eigen:
            multi.check_backend(ctx)
        if eigen and ctx.rank != 0:
            K = None  # only rank 0 decomposes (multi.setup_eigen broadcasts U and d): no float64 copy of K elsewhere
        else:
            K = np.asarray(K, dtype=np.float64)
            if Z is not None:
                Z = np.asarray(Z, dtype=np.float64)
                K = Z @ K @ Z.T  # lmm/lmm.py:124-125
            if eigen and K.shape != (n, n):
                raise ValueError(f"K must be ({n}, {n}) when eigen=True, got {K.shape}")
            if not eigen and K.reshape(-1).shape[0] != n:
                raise ValueError(f"with eigen=False K must be the ({n},) eigenvalue vector, got {K.shape}")

    if not disable_checks:
        # lmm/lmm.py:253-256 tests the rotated arrays; a NaN survives (and only arises from) the rotation
        # of a NaN input, so the inputs are tested instead of downloading U^T X
        if (X.dtype.kind == "f" and np.isnan(X).any()) or np.isnan(Y).any() or np.isnan(W).any():
            raise ValueError("NaNs present in data")

    timing = {"n": n, "m": m, "c0": c0, "q": q, "world_size": ctx.world_size}
    h = factor.handle(c0) if factor is not None else _capi.Handle(n, c0, ctx.device)
    try:
        t0 = time.time()
        if factor is not None:
            timing["eig_ms"] = 0.0  # done once in lmm.factorize
        elif eigen:
            _log(verbose, "Starting eigendecomposition...")
            timing["eig_ms"] = multi.setup_eigen(ctx, h, K)
            _log(verbose, f"Eigendecomposition computed - {round(time.time() - t0, 3)} s")
        else:
            h.set_eigen(None, K.reshape(-1))
        h.set_scan_mode(_capi.PG_SCAN_DE if de else _capi.PG_SCAN_WALD)
        a, b = ctx.shard(m)  # contiguous SNP range of this rank (SampleIter, lmm/lmm.py:427-434)
        outs = []
        nulls = []
        timing["design_ms"], timing["scan_wall_s"], timing["gather_s"] = 0.0, 0.0, 0.0
        res = None
        # traits go through in groups of PG_MAX_TRAITS per pass over the genotypes (one group for lmm.pygemma)
        for q0 in range(0, q, _capi.PG_MAX_TRAITS):
            Yg = Y[:, q0:q0 + _capi.PG_MAX_TRAITS]
            qg = Yg.shape[1]
            t0 = time.time()
            timing["design_ms"] += h.set_design(W, Yg if qg > 1 else Yg.reshape(-1), already_rotated=not eigen)
            _log(verbose, f"Rotated Y, W and built lambda tables - {round(time.time() - t0, 3)} s")
            _log(verbose, f"Running {m} SNPs with {n} individuals...")
            t0 = time.time()
            res = h.scan(X[:, a:b], grid=grid, lrt=lrt)
            if lrt:
                nulls.extend(h.null_model(ph) for ph in range(qg))
            timing["scan"] = res["timing"]
            t1 = time.time()
            if qg == 1:
                outs.append(multi.gather_results(ctx, res, m))
            else:
                outs.extend(multi.gather_results(ctx, {k: v[ph] for k, v in res.items() if k != "timing"}, m)
                            for ph in range(qg))
            timing["gather_s"] += time.time() - t1
            timing["scan_wall_s"] += time.time() - t0
        if res is not None:
            _log(verbose, f"Finished testing {m} SNPs in {round(timing['scan_wall_s'], 3)} s "
                          f"(device: rotate {res['timing']['rotate_ms']:.1f} ms, REML {res['timing']['reml_ms']:.1f} ms)")
    finally:
        if factor is None:
            h.close()

    frames = []
    for out in outs:
        bad = (out["status"] & 1) != 0   # bit 0 = failed row; higher bits are notes (include/pygemma_b200.h)
        data = {}
        for c in COLUMNS:
            col = out[c]
            if bad.any():
                col = col.copy()
                col[bad] = np.nan  # lmm/lmm.py:484-493
            data[c] = col
        cols = list(COLUMNS)
        if lrt:
            for c, k in zip(LRT_COLUMNS, ("D_lrt", "p_lrt", "loglik_ml")):
                col = out[k]
                if bad.any():
                    col = col.copy()
                    col[bad] = np.nan
                data[c] = col
            cols += LRT_COLUMNS
        results_df = pd.DataFrame(data, columns=cols)
        if snps is not None:
            results_df["SNPs"] = snps  # lmm/lmm.py:408-409 (same statement: pandas alignment semantics preserved)
        frames.append(results_df)
    timing["total_s"] = time.time() - t_start
    timing["n_eval2_mean"] = float(np.mean([o["n_eval2"].mean() for o in outs])) if m else 0.0
    timing["n_eval3_mean"] = float(np.mean([o["n_eval3"].mean() for o in outs])) if m else 0.0
    last_timing = timing
    if lrt:
        last_null_model = nulls
    return frames
