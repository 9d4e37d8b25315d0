"""One int8 rotation with the fused tcgen05 engine for ncu: python tools/prof_tc.py n m [fuse|plain] [c0]  (fuse: moments fused in)"""
import sys
import numpy as np
sys.path.insert(0, ".")
from pygemma_b200 import _capi
from pygemma_b200.synth import make_spectral_problem
n, m, c0 = int(sys.argv[1]), int(sys.argv[2]), (int(sys.argv[4]) if len(sys.argv) > 4 else 4)
p = make_spectral_problem(n, m, c0, seed=1, xdtype=np.float64)
rng = np.random.default_rng(0)
U = rng.standard_normal((n, n)).astype(np.float64)
X8 = rng.integers(0, 3, size=(n, m), dtype=np.int8)
with _capi.Handle(n, c0) as h:
    h.set_eigen(U, np.sort(np.abs(p["d"])))
    h.set_design(p["W"], p["Y"])
    h.set_options(rotation=_capi.PG_ROT_I8TC, block_snps=m)
    if len(sys.argv) > 3 and sys.argv[3] == "fuse":
        h.set_moment_fusion(1)
    for rep in range(2):
        o = h.scan(X8)
    print(o["timing"])
