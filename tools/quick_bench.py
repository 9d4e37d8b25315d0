"""Quick device-timing probe used during development (not the contract bench: see bench.py)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from pygemma_b200 import _capi
from pygemma_b200.synth import make_spectral_problem


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    c0 = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    p = make_spectral_problem(n, m, c0, seed=1, xdtype=np.float64)
    out = {"n": n, "m": m, "c0": c0}
    with _capi.Handle(n, c0) as h:
        h.set_eigen(None, p["d"])
        out["design_ms"] = h.set_design(p["W"], p["Y"], already_rotated=True)
        for grid in (False, True):
            for rep in range(2):
                t = time.time()
                o = h.scan(p["X"], grid=grid)
                wall = time.time() - t
            tm = o["timing"]
            out["grid" if grid else "default"] = {
                "reml_ms": tm["reml_ms"], "compress_ms": tm["compress_ms"], "n_nodes": tm["n_nodes"], "snps_per_s_reml": m / (tm["reml_ms"] * 1e-3), "total_ms": tm["total_ms"],
                "wall_s": wall, "ev2": float(o["n_eval2"].mean()), "ev3": float(o["n_eval3"].mean()),
                "status_bad": int((o["status"] != 0).sum())}
    # rotation timing with a random (non-orthogonal) U: timing only
    rng = np.random.default_rng(0)
    U = rng.standard_normal((n, n))
    X8 = rng.integers(0, 3, size=(n, m), dtype=np.int8)
    with _capi.Handle(n, c0) as h:
        h.set_eigen(U, np.abs(p["d"]))
        h.set_design(p["W"], p["Y"])
        for rep in range(2):
            o = h.scan(X8)
        tm = o["timing"]
        out["rotate_int8"] = {k: tm[k] for k in ("total_ms", "h2d_ms", "convert_ms", "rotate_ms", "reml_ms", "compress_ms", "d2h_ms", "n_blocks", "n_nodes")}
        out["rotate_int8"]["snps_per_s_total"] = m / (tm["total_ms"] * 1e-3)
        out["rotate_int8"]["dgemm_tflops"] = 2.0 * n * n * m / (tm["rotate_ms"] * 1e-3) / 1e12
    print(json.dumps(out))


if __name__ == "__main__":
    main()
