"""Development probe: timing + bit check of PG_ROT_I8TC at n = 10 000."""
import sys
import numpy as np
sys.path.insert(0, ".")
from pygemma_b200 import _capi
from pygemma_b200.synth import make_spectral_problem
n, m, c0 = 10000, int(sys.argv[1]) if len(sys.argv) > 1 else 25088, 4
rng = np.random.default_rng(7)
p = make_spectral_problem(n, m, c0, seed=1, xdtype=np.float64)
U = rng.standard_normal((n, n)); X8 = rng.integers(0, 3, size=(n, m), dtype=np.int8)
res = {}
with _capi.Handle(n, c0) as h:
    h.set_eigen(U, np.sort(np.abs(p["d"]))); h.set_design(p["W"], p["Y"])
    for eng in (_capi.PG_ROT_I8SPLIT, _capi.PG_ROT_I8TC):
        h.set_options(rotation=eng, block_snps=m)
        for rep in range(3):
            o = h.scan(X8)
        xr, _ = h.probe_rotated(256)
        res[eng] = xr
        t = o["timing"]
        print("engine", eng, "rotate_ms", round(t["rotate_ms"], 3), "convert_ms", round(t["convert_ms"], 3), "Pop/s", round(7 * 2 * n * n * m / (t["rotate_ms"] * 1e-3) / 1e15, 3), flush=True)
print("bits identical:", np.array_equal(res[_capi.PG_ROT_I8SPLIT], res[_capi.PG_ROT_I8TC]))
