"""Device GRM timing: python tools/grm_bench.py [n] [p]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from pygemma_b200 import _capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
rng = np.random.default_rng(0)
G = rng.binomial(2, 0.3, size=(n, p)).astype(np.int8)
with _capi.Handle(n, 1) as h:
    for rep in range(2):
        t = time.time(); r = h.grm(G, return_K=False, set_kinship=True); w = time.time() - t
    print({"n": n, "p": p, "grm_ms": r["grm_ms"], "eig_ms": r["eig_ms"], "wall_s": w, "tflops_fp64": n * n * p / (r["grm_ms"] * 1e-3) / 1e12})
