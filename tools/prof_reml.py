"""One REML-only scan (rotated inputs) for ncu: python tools/prof_reml.py n m c0 [grid]"""
import sys

import numpy as np

sys.path.insert(0, ".")
from pygemma_b200 import _capi
from pygemma_b200.synth import make_spectral_problem

n, m, c0 = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
grid = len(sys.argv) > 4 and sys.argv[4] == "grid"
p = make_spectral_problem(n, m, c0, seed=1, xdtype=np.float64)
with _capi.Handle(n, c0) as h:
    h.set_eigen(None, p["d"])
    h.set_design(p["W"], p["Y"], already_rotated=True)
    h.set_options(block_snps=m)
    for rep in range(2):
        o = h.scan(p["X"], grid=grid)
    print(o["timing"], float(o["n_eval2"].mean()), float(o["n_eval3"].mean()))
