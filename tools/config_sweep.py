"""Throughput of the BASELINE.json parity configs (shape-matched synthetic data), device timings per stage.
   python tools/config_sweep.py [c1|c2|c3|c5|c4 ...]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch

from pygemma_b200 import _capi

CONFIGS = {
    "c1": dict(n=1940, m=12226, c0=11, grid=False),     # mouse_hs1940 shape
    "c2": dict(n=449, m=100000, c0=6, grid=False),      # GD449 shape
    "c3": dict(n=10000, m=100000, c0=10, grid=False),
    "c5": dict(n=10000, m=100000, c0=40, grid=True),
    "c4": dict(n=50000, m=16384, c0=10, grid=False),    # 20 GB fp64 U; bounded m
}


def make(n, m, c0, seed=3):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(seed)
    if n <= 12000:
        mk = 2 * n
        maf = torch.rand(mk, generator=g, device=dev) * 0.45 + 0.05
        G = ((torch.rand(n, mk, generator=g, device=dev) < maf).double() + (torch.rand(n, mk, generator=g, device=dev) < maf).double())
        sd = G.std(dim=0); sd[sd == 0] = 1.0
        G = (G - G.mean(dim=0)) / sd
        K = (G @ G.T / mk); K.diagonal().add_(1e-3)
        u = G @ torch.randn(mk, generator=g, device=dev, dtype=torch.float64) / mk ** 0.5
        del G
        eig = ("K", K.cpu().numpy())
    else:
        # too large for a benchmark-time syevd: random orthogonal U (QR of a gaussian) and a chi-square-like spectrum
        A = torch.randn(n, n, generator=g, device=dev, dtype=torch.float64)
        Q, _ = torch.linalg.qr(A)
        del A
        d = (torch.rand(n, generator=g, device=dev, dtype=torch.float64) * 3 + 0.05).sort().values
        u = Q @ (d.sqrt() * torch.randn(n, generator=g, device=dev, dtype=torch.float64))
        eig = ("U", Q, d)
    u = u / u.std()
    W = torch.cat([torch.ones(n, 1, device=dev, dtype=torch.float64), torch.randn(n, c0 - 1, generator=g, device=dev, dtype=torch.float64)], dim=1)
    y = 0.7 * u + 0.7 * torch.randn(n, generator=g, device=dev, dtype=torch.float64) + 0.05 * W[:, 1:].sum(dim=1)
    X = torch.empty((n, m), dtype=torch.int8, device=dev)
    for a in range(0, m, 8192):
        b = min(m, a + 8192)
        mf = torch.rand(b - a, generator=g, device=dev) * 0.45 + 0.05
        X[:, a:b] = ((torch.rand(n, b - a, generator=g, device=dev) < mf).to(torch.int8) + (torch.rand(n, b - a, generator=g, device=dev) < mf).to(torch.int8))
    return eig, W.cpu().numpy(), y.cpu().numpy(), X


for name in (sys.argv[1:] or ["c1", "c2", "c5"]):
    cfg = CONFIGS[name]
    n, m, c0, grid = cfg["n"], cfg["m"], cfg["c0"], cfg["grid"]
    eig, W, y, X = make(n, m, c0)
    out = torch.empty((6, m), dtype=torch.float64, device="cuda:0")
    st = torch.zeros((3, m), dtype=torch.int32, device="cuda:0")
    with _capi.Handle(n, c0) as h:
        t0 = time.time()
        if eig[0] == "K":
            _, eig_ms = h.set_kinship(eig[1])
        else:
            h.set_eigen_device(eig[1].t().contiguous().data_ptr(), False, eig[2].data_ptr())  # column-major U
            eig_ms = 0.0
        setup_s = time.time() - t0
        design_ms = h.set_design(W, y)
        for rep in range(3):
            tm = h.scan_device(X.data_ptr(), _capi.PG_X_I8, m, _capi.PG_X_SAMPLE_MAJOR, m, grid, [out[i].data_ptr() for i in range(6)],
                               st[0].data_ptr(), st[1].data_ptr(), st[2].data_ptr())
    o = out.cpu().numpy()
    res = {"config": name, **cfg, "snps_per_s": m / (tm["total_ms"] * 1e-3), "eig_ms": eig_ms, "setup_s": setup_s, "design_ms": design_ms,
           "timing": {k: (round(v, 3) if isinstance(v, float) else v) for k, v in tm.items()},
           "nan_rows": int(np.isnan(o).any(axis=0).sum()), "bad_status": int((st[0] != 0).sum()),
           "ev2": float(st[1].float().mean()), "ev3": float(st[2].float().mean()),
           "lambda_quantiles": [float(q) for q in np.nanquantile(o[3], [0.01, 0.5, 0.99])]}
    print(json.dumps(res), flush=True)
    del X, out, st, eig
    torch.cuda.empty_cache()
