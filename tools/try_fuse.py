"""Development probe: moments fused into the rotation (pg_set_moment_fusion) against the unfused path and the oracle."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import oracle
from pygemma_b200 import _capi

COLS = ["beta", "se_beta", "tau", "lambda", "F_wald", "p_wald"]


def problem(n, m, c0, seed, span=1.5):
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.standard_normal((n, n)))
    d = np.sort(10.0 ** (rng.random(n) * span - 1.0))
    d[: max(1, n // 100)] = 0.0                      # a null space
    d[-3:] *= np.array([3.0, 10.0, 40.0])           # isolated top eigenvalues (COPY rows)
    W = np.column_stack([np.ones(n)] + [rng.standard_normal(n) for _ in range(c0 - 1)])
    u = U @ (np.sqrt(d) * rng.standard_normal(n))
    y = 0.7 * u / u.std() + 0.7 * rng.standard_normal(n) + 0.05 * W[:, 1:].sum(1)
    maf = rng.random(m) * 0.45 + 0.05
    X = ((rng.random((n, m)) < maf).astype(np.int8) + (rng.random((n, m)) < maf).astype(np.int8))
    return U, d, W, y, X


def rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def case(n, m, c0, seed, q=1, xdtype=np.int8):
    U, d, W, y, X = problem(n, m, c0, seed)
    X = X.astype(xdtype)
    out = {"n": n, "m": m, "c0": c0}
    with _capi.Handle(n, c0) as h:
        h.set_eigen(U, d)
        h.set_design(W, y)
        res = {}
        for mode in (0, 1):
            h.set_moment_fusion(mode)
            o = h.scan(X)
            res[mode] = o
            out[f"engine_{mode}"] = o["timing"]["rot_engine"]
            out[f"info_{mode}"] = h.fusion_info()
        out["nodes"] = res[0]["timing"]["n_nodes"]
        out["fused_vs_unfused"] = {c: float(np.nanmax(rel(res[1][c], res[0][c]))) for c in COLS}
        out["status_equal"] = bool(np.array_equal(res[0]["status"], res[1]["status"]))
    idx = np.unique(np.linspace(0, m - 1, 48).astype(np.int64))
    xr = U.T @ X[:, idx].astype(np.float64)
    ref = oracle.scan_rotated(d, U.T @ y, U.T @ W, np.ascontiguousarray(xr.T))
    for mode in (0, 1):
        out[f"vs_oracle_{mode}"] = {c: float(np.nanmax(rel(res[mode][c][idx], ref[c]))) for c in COLS}
    return out


if __name__ == "__main__":
    cases = [(1024, 1500, 5, 1), (1000, 777, 3, 2), (2048, 4096 + 300, 10, 3)]
    if len(sys.argv) > 1:
        cases = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
    for cs in cases:
        print(json.dumps(case(*cs)), flush=True)
