"""Development probe: fused tcgen05 rotation (PG_ROT_I8TC) vs cuBLAS int8 split (PG_ROT_I8SPLIT): bits and time."""
import sys
import numpy as np
sys.path.insert(0, ".")
from pygemma_b200 import _capi
from pygemma_b200.synth import make_spectral_problem

def run(n, m, c0=4, blk=0):
    rng = np.random.default_rng(n + m)
    p = make_spectral_problem(n, m, c0, seed=1, xdtype=np.float64)
    U = rng.standard_normal((n, n))
    X8 = rng.integers(0, 3, size=(n, m), dtype=np.int8)
    res = {}
    with _capi.Handle(n, c0) as h:
        h.set_eigen(U, np.sort(np.abs(p["d"])))
        h.set_design(p["W"], p["Y"])
        for eng in (_capi.PG_ROT_I8SPLIT, _capi.PG_ROT_I8TC):
            h.set_options(rotation=eng, block_snps=blk)
            for rep in range(2):
                o = h.scan(X8)
            cnt = min(m, 512)
            xr, row0 = h.probe_rotated(cnt)
            res[eng] = (o, xr, row0)
            print(n, m, "engine", eng, "rotate_ms", round(o["timing"]["rotate_ms"], 3), "convert_ms", round(o["timing"]["convert_ms"], 3),
                  "blocks", o["timing"]["n_blocks"], flush=True)
    a, b = res[_capi.PG_ROT_I8SPLIT], res[_capi.PG_ROT_I8TC]
    same = np.array_equal(a[1], b[1])
    print("  rotated bits identical:", same, "max abs diff", float(np.abs(a[1] - b[1]).max()), "max |x|", float(np.abs(a[1]).max()))
    if not same:
        bad = np.argwhere(a[1] != b[1])
        print("  first mismatches (snp, eig):", bad[:8].tolist(), "count", len(bad), "of", a[1].size)
        i, j = bad[0]
        print("  values", a[1][i, j], b[1][i, j])
    print("  outputs identical:", all(np.array_equal(a[0][c], b[0][c], equal_nan=True) for c in ("beta", "lambda", "p_wald")))

for (n, m) in [(256, 256), (777, 300), (2048, 1024), (10000, 25088)]:
    run(n, m)
