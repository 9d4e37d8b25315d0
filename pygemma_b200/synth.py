"""Seeded synthetic GWAS inputs shaped like the reference's benchmark data.

The reference's own datasets are not bundled (its genotype and kinship files
are missing blobs), so tests and bench.py use shape-matched synthetic inputs:
binomial dosages, a kinship matrix from an independent standardised panel,
an intercept + Gaussian covariates, and a phenotype with a polygenic part.
Mirrors how the reference's callers prepare inputs
(reference experiments/benchmarks/benchmarks.py:233-313,
experiments/wtccc/run_pygemma.py:410-447).
"""
from __future__ import annotations

import numpy as np


def make_genotypes(n: int, m: int, rng: np.random.Generator, dtype=np.int8) -> np.ndarray:
    """(n, m) sample-major dosage matrix X_ij ~ Binomial(2, MAF_j), MAF_j ~ U(0.05, 0.5)."""
    maf = rng.uniform(0.05, 0.5, size=m)
    out = np.empty((n, m), dtype=dtype)
    step = max(1, (1 << 24) // max(n, 1))
    for a in range(0, m, step):
        b = min(m, a + step)
        out[:, a:b] = rng.binomial(2, maf[a:b][None, :], size=(n, b - a)).astype(dtype)
    return out


def make_kinship(n: int, rng: np.random.Generator, m_k: int | None = None, ridge: float = 1e-3):
    """K = G G^T / m_k + ridge*I from an independent standardised binomial panel; returns (K, G)."""
    m_k = m_k or max(2000, 2 * n)
    g = make_genotypes(n, m_k, rng, dtype=np.float64)
    sd = g.std(axis=0)
    sd[sd == 0] = 1.0
    g = (g - g.mean(axis=0)) / sd
    k = g @ g.T / m_k
    k[np.diag_indices(n)] += ridge
    return k, g


def make_problem(n: int, m: int, c0: int, seed: int = 0, h2: float = 0.5, n_causal: int = 5,
                 xdtype=np.int8, m_k: int | None = None):
    """Returns dict(Y (n,1) f64, X (n,m) xdtype, W (n,c0) f64 with intercept first, K (n,n) f64)."""
    rng = np.random.default_rng(seed)
    x = make_genotypes(n, m, rng, dtype=np.int8)
    k, g = make_kinship(n, rng, m_k=m_k)
    w = np.concatenate([np.ones((n, 1)), rng.standard_normal((n, max(c0 - 1, 0)))], axis=1)[:, :c0]
    u = g @ rng.standard_normal(g.shape[1]) / np.sqrt(g.shape[1])
    u = u / (u.std() + 1e-300)
    e = rng.standard_normal(n)
    y = np.sqrt(h2) * u + np.sqrt(max(1.0 - h2, 0.0)) * e
    if n_causal and m:
        idx = rng.choice(m, size=min(n_causal, m), replace=False)
        xs = x[:, idx].astype(np.float64)
        sd = xs.std(axis=0)
        sd[sd == 0] = 1.0
        y = y + 0.1 * ((xs - xs.mean(axis=0)) / sd).sum(axis=1)
    y = y + w[:, 1:].sum(axis=1) * 0.05
    return {"Y": y.reshape(-1, 1), "X": x.astype(xdtype), "W": w, "K": k}


def make_spectral_problem(n: int, m: int, c0: int, seed: int = 0, h2: float = 0.5, xdtype=np.int8):
    """Large-n variant with no dense K: returns rotated-space inputs for the eigen=False entry.

    d ~ scaled chi-square-like spectrum; Y, W, X are drawn directly in the
    rotated basis (X gaussian with per-SNP scale, as U^T x would be).  Used for
    REML-kernel benchmarks where forming K and running syevd would dominate.
    """
    rng = np.random.default_rng(seed)
    d = np.sort(rng.chisquare(4, size=n) / 4.0)
    w = np.concatenate([np.ones((n, 1)), rng.standard_normal((n, max(c0 - 1, 0)))], axis=1)[:, :c0]
    y = np.sqrt(h2 * d + (1 - h2)) * rng.standard_normal(n)
    x = rng.standard_normal((n, m)) * rng.uniform(0.3, 0.7, size=m)[None, :]
    if np.issubdtype(np.dtype(xdtype), np.integer):
        x = np.clip(np.rint(x + 1.0), 0, 2)
    return {"Y": y.reshape(-1, 1), "X": x.astype(xdtype), "W": w, "d": d}
