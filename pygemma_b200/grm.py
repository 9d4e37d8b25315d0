"""Genetic relatedness matrix on the GPU -- the step the reference's callers run right before `lmm.pygemma`.

    from pygemma_b200 import grm
    K = grm.calculate_genetic_relatedness_matrix(X)          # experiments/animal_gwas/run_gwas.py:45-55

Same definition as the reference's helper (and as `K = X @ X.T / p` on StandardScaler output,
experiments/wtccc/run_pygemma.py:432,:447): columns are centred and divided by their population standard
deviation (constant columns: divisor 1), K = Z Z^T / p.  The work runs in libpygemma_b200.so (pg_grm): SNP blocks
are standardised on the device and accumulated with an FP64 SYRK; there is no CPU fallback.
"""
from __future__ import annotations

import numpy as np

from . import _capi

last_timing: dict = {}


def calculate_genetic_relatedness_matrix(X, device: int = 0):
    """X: (n, p) marker matrix (int8 dosages, float32 or float64).  Returns the (n, n) float64 GRM."""
    global last_timing
    X = np.asarray(X)
    if X.ndim != 2:
        raise ValueError("X must have shape (n, p)")
    if X.dtype not in (np.int8, np.float32, np.float64):
        X = X.astype(np.int8) if (np.issubdtype(X.dtype, np.integer) and X.size and X.min() >= -128 and X.max() <= 127) \
            else X.astype(np.float64)
    n = X.shape[0]
    with _capi.Handle(n, 1, device) as h:
        r = h.grm(X)
    last_timing = {"grm_ms": r["grm_ms"], "n": n, "p": X.shape[1]}
    return r["K"]
