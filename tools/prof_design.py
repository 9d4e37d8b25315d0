"""pg_set_design a few times for a launch list: python tools/prof_design.py n c0"""
import sys
import numpy as np
sys.path.insert(0, ".")
from pygemma_b200 import _capi
from pygemma_b200.synth import make_spectral_problem
n, c0 = int(sys.argv[1]), int(sys.argv[2])
p = make_spectral_problem(n, 8, c0, seed=1, xdtype=np.float64)
rng = np.random.default_rng(0)
U = np.linalg.qr(rng.standard_normal((n, n)))[0] if n <= 4000 else rng.standard_normal((n, n))
with _capi.Handle(n, c0) as h:
    h.set_eigen(U, p["d"])
    for rep in range(3):
        print(h.set_design(p["W"], p["Y"]))
