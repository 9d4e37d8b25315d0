"""Throughput of the packed PLINK .bed path at n = 10 000: python tools/bed_bench.py [m] [missing_rate]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from pygemma_b200 import _capi
from pygemma_b200.synth import make_spectral_problem
n, c0 = 10000, 10
m = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
rng = np.random.default_rng(3)
p = make_spectral_problem(n, 8, c0, seed=1, xdtype=np.float64)
U = rng.standard_normal((n, n))
res = {}
with _capi.Handle(n, c0) as h:
    h.set_eigen(U, np.sort(np.abs(p["d"]))); h.set_design(p["W"], p["Y"])
    for miss in (0.0, 0.02):
        two = rng.choice(np.array([0, 2, 3], dtype=np.uint8), size=(m, n), p=[0.5, 0.4, 0.1])
        if miss:
            two[rng.random((m, n)) < miss] = 1
        two = two.reshape(m, n // 4, 4)
        packed = (two[:, :, 0] | (two[:, :, 1] << 2) | (two[:, :, 2] << 4) | (two[:, :, 3] << 6)).astype(np.uint8)
        del two
        for rep in range(3):
            t = time.time(); o = h.scan_bed(packed); w = time.time() - t
        tm = o["timing"]
        print({"missing_rate": miss, "m": m, "wall_ms": round(w * 1e3, 1), "total_ms": round(tm["total_ms"], 1), "h2d_ms": round(tm["h2d_ms"], 1),
               "convert_ms": round(tm["convert_ms"], 1), "rotate_ms": round(tm["rotate_ms"], 1), "reml_ms": round(tm["reml_ms"], 1),
               "snps_per_s_wall": round(m / w), "nan_rows": int(np.isnan(o["beta"]).sum())}, flush=True)
