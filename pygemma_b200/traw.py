"""PLINK `.traw` ingest and the callers' `std > 0` filter.

`.traw` is PLINK's variant-major additive recoding (`--recode A-transpose`): a tab-separated text table with the columns
`CHR SNP (C)M POS COUNTED ALT` followed by one column per sample (`FID_IID`) holding the count of the COUNTED allele
(0 / 1 / 2) or `NA`.  The reference bundles its 1000G example in this format (data/CCDG_TGP.GD449...100k.traw.gz, listed
in .MISSING_LARGE_BLOBS) and BASELINE.json configs[1] is quoted on it.  Parsing text is host work; what goes to the GPU is
the same packed 2-bit block the `.bed` path uploads (pygemma_b200.bed), so missing calls are mean-imputed, columns
optionally standardised and dosages rotated on the exact int8 tensor-core path without ever existing as floats.

The reference's callers drop monomorphic markers before the scan (experiments/wtccc/run_pygemma.py:407-410:
`X_std = X.std(axis=0); snp_info = snp_info[X_std > 0]; X = X[:, X_std > 0]`); `std_filter` reproduces that mask from
the packed genotype counts, `pygemma_traw(..., filter_std=True)` applies it and returns it.

    from pygemma_b200 import traw
    df, keep = traw.pygemma_traw(Y, "chr20.traw.gz", W, K, filter_std=True)
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from . import _capi, bed

INFO_COLUMNS = ["CHR", "SNP", "(C)M", "POS", "COUNTED", "ALT"]


def read_traw(path: str, chunk_snps: int = 8192):
    """Returns (packed, n, info, samples): packed is the (m, ceil(n/4)) uint8 block in PLINK .bed coding of the COUNTED
    allele counts (count_A1=False decoding gives them back: 00 -> 0, 10 -> 1, 11 -> 2, 01 -> missing), info the six
    leading columns as a DataFrame, samples the sample column names.  Reads in chunks of `chunk_snps` rows; `.gz` /
    `.bz2` / `.xz` are decompressed by pandas."""
    head = pd.read_csv(path, sep="\t", nrows=0)
    cols = list(head.columns)
    if len(cols) < 7 or [c.upper() for c in cols[:2]] != ["CHR", "SNP"]:
        raise ValueError(f"{path}: not a PLINK .traw table (expected CHR SNP (C)M POS COUNTED ALT <samples...>)")
    samples = cols[6:]
    n = len(samples)
    dtypes = {c: np.float32 for c in samples}
    dtypes.update({c: str for c in cols[:6]})
    packed, infos = [], []
    for chunk in pd.read_csv(path, sep="\t", na_values=["NA", "nan", "."], keep_default_na=True, dtype=dtypes,
                             chunksize=chunk_snps):
        infos.append(chunk[cols[:6]])
        g = chunk[samples].to_numpy(dtype=np.float64)   # (rows = SNPs, n)
        bad = ~(np.isnan(g) | (g == 0.0) | (g == 1.0) | (g == 2.0))
        if bad.any():
            r, c = np.argwhere(bad)[0]
            raise ValueError(f"{path}: allele count {g[r, c]!r} for SNP {chunk.iloc[r, 1]} / sample {samples[c]}: "
                             ".traw holds 0, 1, 2 or NA (real-valued dosages go through lmm.pygemma as floats)")
        packed.append(bed.encode_packed(g.T))
    info = pd.concat(infos, ignore_index=True) if infos else head[cols[:6]]
    info.columns = INFO_COLUMNS
    block = np.concatenate(packed, axis=0) if packed else np.zeros((0, (n + 3) // 4), dtype=np.uint8)
    return block, n, info, samples


def write_traw(path: str, G: np.ndarray, snp_ids=None, samples=None) -> None:
    """Writes an (n, m) matrix of allele counts in {0, 1, 2, NaN} as a .traw table (fixtures for the tests)."""
    n, m = G.shape
    ids = list(snp_ids) if snp_ids is not None else [f"rs{i}" for i in range(m)]
    samples = list(samples) if samples is not None else [f"f{j}_i{j}" for j in range(n)]
    body = pd.DataFrame(np.asarray(G, dtype=np.float64).T, columns=samples)
    body = body.apply(lambda col: col.map(lambda v: "NA" if np.isnan(v) else str(int(v))))
    info = pd.DataFrame({"CHR": ["1"] * m, "SNP": ids, "(C)M": ["0"] * m, "POS": [str(i + 1) for i in range(m)],
                         "COUNTED": ["A"] * m, "ALT": ["G"] * m})
    pd.concat([info, body], axis=1).to_csv(path, sep="\t", index=False)


_POP = np.array([[(b >> (2 * i)) & 3 for i in range(4)] for b in range(256)], dtype=np.uint8)   # four 2-bit codes per byte


def genotype_counts(packed: np.ndarray, n: int) -> np.ndarray:
    """(m, 4) counts of the codes 00 (count 0), 01 (missing), 10 (count 1), 11 (count 2) among the n samples of each SNP."""
    packed = np.asarray(packed, dtype=np.uint8)
    m, bps = packed.shape
    counts = np.zeros((m, 4), dtype=np.int64)
    per_byte = np.stack([(_POP == c).sum(axis=1) for c in range(4)], axis=1)   # (256, 4)
    full = n // 4
    if full:
        counts += per_byte[packed[:, :full]].sum(axis=1)
    for i in range(n - 4 * full):   # ragged last byte: padding bits are not samples
        code = (packed[:, full] >> (2 * i)) & 3
        counts[np.arange(m), code] += 1
    return counts


def std_filter(packed: np.ndarray, n: int, nan_policy: str = "propagate") -> np.ndarray:
    """The callers' monomorphic-marker filter `X.std(axis=0) > 0` (experiments/wtccc/run_pygemma.py:407-410) as a boolean
    keep-mask over SNPs.  nan_policy='propagate' is NumPy's (and hence the reference's) behaviour on pysnptools output:
    a column with a missing call has std NaN and NaN > 0 is False, so it is dropped; 'omit' judges the observed calls."""
    if nan_policy not in ("propagate", "omit"):
        raise ValueError("nan_policy must be 'propagate' or 'omit'")
    c = genotype_counts(packed, n)
    classes = (c[:, [0, 2, 3]] > 0).sum(axis=1)
    keep = classes >= 2
    if nan_policy == "propagate":
        keep &= c[:, 1] == 0
    return keep


def pygemma_traw(Y, path: str, W, K, filter_std: bool = False, nan_policy: str = "propagate", standardize: bool = False,
                 grid: bool = False, verbose: int = 0, device: int = 0):
    """lmm.pygemma for the genotypes of a `.traw` file.  Returns (DataFrame, keep): the reference's six columns plus
    'SNPs' (the .traw SNP ids) for the kept markers in file order, and the boolean keep-mask over all markers of the
    file (all True without filter_std).  Missing calls are mean-imputed (SimpleImputer(strategy='mean'),
    experiments/benchmarks/benchmarks.py:22), `standardize` applies StandardScaler semantics -- both on the device."""
    packed, n, info, _ = read_traw(path)
    keep = std_filter(packed, n, nan_policy) if filter_std else np.ones(packed.shape[0], dtype=bool)
    Y = np.asarray(Y, dtype=np.float64).reshape(-1)
    W = np.asarray(W, dtype=np.float64)
    if W.ndim == 1:
        W = W.reshape(-1, 1)
    if Y.shape[0] != n or W.shape[0] != n:
        raise ValueError(f"{path} lists {n} samples; Y has {Y.shape[0]}, W has {W.shape[0]}")
    sel = np.ascontiguousarray(packed[keep])
    with _capi.Handle(n, W.shape[1], device) as h:
        h.set_kinship(np.asarray(K, dtype=np.float64))
        h.set_design(W, Y)
        out = h.scan_bed(sel, grid=grid, count_A1=False, standardize=standardize)
    if verbose > 0:
        print(f"[pygemma_b200] {sel.shape[0]} of {packed.shape[0]} SNPs from {path}: {out['timing']}", flush=True)
    bad = (out["status"] & 1) != 0
    data = {}
    for c in ("beta", "se_beta", "tau", "lambda", "F_wald", "p_wald"):
        col = out[c]
        if bad.any():
            col = col.copy()
            col[bad] = np.nan
        data[c] = col
    df = pd.DataFrame(data)
    df["SNPs"] = info["SNP"].to_numpy()[keep]
    return df, keep
