"""Development probe: pg_scan from pageable host memory (what lmm.pygemma callers pass), with / without the bounce path."""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
from pygemma_b200 import _capi
from pygemma_b200.synth import make_spectral_problem
n, m, c0 = 10000, 100000, 10
rng = np.random.default_rng(7)
p = make_spectral_problem(n, 8, c0, seed=1, xdtype=np.float64)
U = rng.standard_normal((n, n)); X8 = rng.integers(0, 3, size=(n, m), dtype=np.int8)
with _capi.Handle(n, c0) as h:
    h.set_eigen(U, np.sort(np.abs(p["d"]))); h.set_design(p["W"], p["Y"])
    for rep in range(3):
        t = time.time(); o = h.scan(X8); w = time.time() - t
        tm = o["timing"]
        print("bounce" if not os.environ.get("PG_NO_BOUNCE") else "pageable 2D copy", "wall_ms", round(w * 1e3, 1), "total_ms", round(tm["total_ms"], 1),
              "h2d_ms", round(tm["h2d_ms"], 1), "SNPs/s (wall)", round(m / w), flush=True)
