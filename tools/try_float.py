"""Development probe: throughput of float64 standardised genotypes (level-coded int8 path) at n = 10 000."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from pygemma_b200 import _capi
from pygemma_b200.synth import make_spectral_problem
n, m, c0 = 10000, 25088, 10
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
maf = torch.rand(m, generator=g, device=dev) * 0.45 + 0.05
G = ((torch.rand(n, m, generator=g, device=dev) < maf).double() + (torch.rand(n, m, generator=g, device=dev) < maf).double())
Z = (G - G.mean(dim=0)) / G.std(dim=0, unbiased=False)
A = torch.randn(n, n, generator=g, device=dev, dtype=torch.float64)
p = make_spectral_problem(n, 8, c0, seed=1, xdtype=np.float64)
out = torch.empty((6, m), dtype=torch.float64, device=dev)
with _capi.Handle(n, c0) as h:
    h.set_eigen_device(A.data_ptr(), False, torch.tensor(np.sort(np.abs(p["d"])), device=dev).data_ptr())
    h.set_design(p["W"], p["Y"])
    for name, X in (("std_f64", Z), ("raw_f64", G), ("std_f32", Z.float())):
        xd = _capi.PG_X_F64 if X.dtype == torch.float64 else _capi.PG_X_F32
        for rep in range(2):
            tm = h.scan_device(X.data_ptr(), xd, m, _capi.PG_X_SAMPLE_MAJOR, m, False, [out[i].data_ptr() for i in range(6)])
        print(name, "engine", tm["rot_engine"], "convert_ms", round(tm["convert_ms"], 2), "rotate_ms", round(tm["rotate_ms"], 2),
              "total_ms", round(tm["total_ms"], 2), "SNPs/s", round(m / tm["total_ms"] * 1e3), flush=True)
    h.set_options(rotation=_capi.PG_ROT_FP64)
    tm = h.scan_device(Z.data_ptr(), _capi.PG_X_F64, m, _capi.PG_X_SAMPLE_MAJOR, m, False, [out[i].data_ptr() for i in range(6)])
    print("std_f64 forced FP64 GEMM: rotate_ms", round(tm["rotate_ms"], 2), "total_ms", round(tm["total_ms"], 2), "SNPs/s", round(m / tm["total_ms"] * 1e3))
