"""Throughput of the multi-phenotype scan (pg_set_design_multi): SNP x trait tests per second against q.
Development probe, not the contract bench (bench.py).  Random (non-orthogonal) U: timing only."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from pygemma_b200 import _capi
from pygemma_b200.synth import make_spectral_problem


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
    c0 = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    qs = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1, 2, 4, 8, 16]
    p = make_spectral_problem(n, 8, c0, seed=1)
    rng = np.random.default_rng(0)
    U = rng.standard_normal((n, n))
    X8 = rng.integers(0, 3, size=(n, m), dtype=np.int8)
    rows = []
    with _capi.Handle(n, c0) as h:
        h.set_eigen(U, np.abs(p["d"]))
        for q in qs:
            Y = rng.standard_normal((n, q))
            t = time.time()
            design_ms = h.set_design(p["W"], Y if q > 1 else Y[:, 0])
            design_wall = time.time() - t
            for rep in range(2):
                t = time.time()
                o = h.scan(X8, with_counts=False)
                wall = time.time() - t
            tm = o["timing"]
            rows.append({"q": q, "design_ms": design_ms, "design_wall_s": design_wall, "total_ms": tm["total_ms"], "rotate_ms": tm["rotate_ms"],
                         "reml_ms": tm["reml_ms"], "compress_ms": tm["compress_ms"], "rot_engine": tm["rot_engine"],
                         "n_nodes": tm["n_nodes"], "wall_s": wall,
                         "snps_per_s": m / (tm["total_ms"] * 1e-3),
                         "tests_per_s": m * q / (tm["total_ms"] * 1e-3),
                         "tests_per_s_wall": m * q / wall,
                         "bad": int((o["status"] != 0).sum())})
            print(json.dumps(rows[-1]), flush=True)
    print(json.dumps({"n": n, "m": m, "c0": c0, "rows": rows}))


if __name__ == "__main__":
    main()
