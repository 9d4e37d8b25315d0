"""PLINK .bed ingest for the scan.

The reference's experiment drivers read `.bed/.bim/.fam` with pysnptools (`Bed(...).read().val`, NaN for missing calls,
experiments/ukb_afr/code/run_snp.py:49-66), impute the column mean (`SimpleImputer(strategy='mean')`,
experiments/benchmarks/benchmarks.py:22, experiments/animal_gwas/run_gwas.py:93-94), optionally standardise, and call
`lmm.pygemma`.  Here the packed file body goes to the GPU as it is (2 bits per genotype): decoding, mean imputation and
standardisation happen on the device and the genotypes enter the exact int8 tensor-core rotation as dosage codes plus a
missing-indicator component -- they never exist as floats.

    from pygemma_b200 import bed
    df = bed.pygemma_bed(Y, "chr20", W, K)                 # chr20.bed / .bim / .fam
    G = bed.read_bed("chr20")                              # (n, m) float64 with NaN, NumPy decode (small files, tests)
"""
from __future__ import annotations

import os

import numpy as np
import pandas as pd

from . import _capi

_MAGIC = bytes([0x6C, 0x1B, 0x01])  # PLINK 1 .bed, SNP-major


def _count_lines(path: str) -> int:
    with open(path, "rb") as f:
        return sum(1 for _ in f)


def read_packed(prefix: str):
    """Returns (packed (m, ceil(n/4)) uint8, n, snp_ids) of `prefix`.bed/.bim/.fam."""
    n = _count_lines(prefix + ".fam")
    bim = pd.read_csv(prefix + ".bim", sep=r"\s+", header=None, usecols=[1], names=["snp"], dtype=str)
    m = len(bim)
    bps = (n + 3) // 4
    with open(prefix + ".bed", "rb") as f:
        if f.read(3) != _MAGIC:
            raise ValueError(f"{prefix}.bed is not a SNP-major PLINK 1 .bed file")
        packed = np.fromfile(f, dtype=np.uint8, count=m * bps)
    if packed.size != m * bps:
        raise ValueError(f"{prefix}.bed is truncated: expected {m} x {bps} bytes")
    return packed.reshape(m, bps), n, bim["snp"].to_numpy()


def decode_packed(packed: np.ndarray, n: int, count_A1: bool = False) -> np.ndarray:
    """NumPy decode of a packed block to an (n, m) float64 dosage matrix with NaN for missing calls
    (what pysnptools' Bed(count_A1=...).read().val returns)."""
    packed = np.asarray(packed, dtype=np.uint8)
    two = np.stack([(packed >> (2 * i)) & 3 for i in range(4)], axis=-1).reshape(packed.shape[0], -1)[:, :n]
    hom_first, hom_second = (2.0, 0.0) if count_A1 else (0.0, 2.0)
    lut = np.array([hom_first, np.nan, 1.0, hom_second])
    return lut[two].T.copy()


def encode_packed(G: np.ndarray, count_A1: bool = False) -> np.ndarray:
    """Inverse of decode_packed for (n, m) dosages in {0, 1, 2, NaN}: a packed (m, ceil(n/4)) block (tests, fixtures)."""
    G = np.asarray(G, dtype=np.float64)
    n, m = G.shape
    code = np.full((m, 4 * ((n + 3) // 4)), 0, dtype=np.uint8)
    gt = G.T
    first, second = (2.0, 0.0) if count_A1 else (0.0, 2.0)
    c = np.where(np.isnan(gt), 1, np.where(gt == 1.0, 2, np.where(gt == second, 3, 0))).astype(np.uint8)
    if not np.all(np.isnan(gt) | (gt == 0.0) | (gt == 1.0) | (gt == 2.0)):
        raise ValueError("dosages must be 0, 1, 2 or NaN")
    del first
    code[:, :n] = c
    code = code.reshape(m, -1, 4)
    return (code[:, :, 0] | (code[:, :, 1] << 2) | (code[:, :, 2] << 4) | (code[:, :, 3] << 6)).astype(np.uint8)


def write_bed(prefix: str, G: np.ndarray, count_A1: bool = False, snp_ids=None) -> None:
    """Writes `prefix`.bed/.bim/.fam for an (n, m) dosage matrix (fixtures for the tests)."""
    n, m = G.shape
    with open(prefix + ".bed", "wb") as f:
        f.write(_MAGIC)
        f.write(encode_packed(G, count_A1).tobytes())
    ids = snp_ids if snp_ids is not None else [f"rs{i}" for i in range(m)]
    with open(prefix + ".bim", "w") as f:
        for i, s in enumerate(ids):
            f.write(f"1\t{s}\t0\t{i + 1}\tA\tG\n")
    with open(prefix + ".fam", "w") as f:
        for j in range(n):
            f.write(f"f{j}\ti{j}\t0\t0\t0\t-9\n")


def read_bed(prefix: str, count_A1: bool = False) -> np.ndarray:
    packed, n, _ = read_packed(prefix)
    return decode_packed(packed, n, count_A1)


def pygemma_bed(Y, prefix: str, W, K, count_A1: bool = False, standardize: bool = False, grid: bool = False,
                eigen: bool = True, verbose: int = 0, device: int = 0):
    """lmm.pygemma for the genotypes of `prefix`.bed: same DataFrame (`beta ... p_wald, SNPs` in file order).
    Missing calls are mean-imputed, `standardize` applies StandardScaler semantics, both on the device."""
    if not eigen:
        raise ValueError("packed genotypes cannot be pre-rotated: pass K (eigen=True)")
    packed, n, ids = read_packed(prefix)
    Y = np.asarray(Y, dtype=np.float64).reshape(-1)
    W = np.asarray(W, dtype=np.float64)
    if W.ndim == 1:
        W = W.reshape(-1, 1)
    if Y.shape[0] != n or W.shape[0] != n:
        raise ValueError(f"{prefix}.fam lists {n} samples; Y has {Y.shape[0]}, W has {W.shape[0]}")
    with _capi.Handle(n, W.shape[1], device) as h:
        h.set_kinship(np.asarray(K, dtype=np.float64))
        h.set_design(W, Y)
        out = h.scan_bed(packed, grid=grid, count_A1=count_A1, standardize=standardize)
    if verbose > 0:
        print(f"[pygemma_b200] {packed.shape[0]} SNPs from {os.path.basename(prefix)}.bed: {out['timing']}", flush=True)
    bad = (out["status"] & 1) != 0
    data = {}
    for c in ("beta", "se_beta", "tau", "lambda", "F_wald", "p_wald"):
        col = out[c]
        if bad.any():
            col = col.copy()
            col[bad] = np.nan
        data[c] = col
    df = pd.DataFrame(data)
    df["SNPs"] = ids
    return df
