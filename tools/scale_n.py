import json, sys, numpy as np
sys.path.insert(0, ".")
from pygemma_b200 import _capi
from pygemma_b200.synth import make_spectral_problem
c0 = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for n, m in ((128, 32768), (1024, 32768), (4096, 32768), (10000, 32768)):
    p = make_spectral_problem(n, m, c0, seed=1, xdtype=np.float64)
    with _capi.Handle(n, c0) as h:
        h.set_eigen(None, p["d"]); h.set_design(p["W"], p["Y"], already_rotated=True)
        h.set_options(block_snps=32768)
        for grid in (False, True):
            for rep in range(2):
                o = h.scan(p["X"], grid=grid)
            print(json.dumps({"n": n, "m": m, "grid": grid, "reml_ms": o["timing"]["reml_ms"], "ev2": float(o["n_eval2"].mean()), "ev3": float(o["n_eval3"].mean()), "us_per_snp_pass": 1e3*o["timing"]["reml_ms"]/m/(o["n_eval2"].mean()+o["n_eval3"].mean())}))
