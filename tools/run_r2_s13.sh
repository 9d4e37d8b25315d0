#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi_device.py tests/test_gpu_fullsize.py -m gpu -q --tb=short -k "two_gpus or multi_phenotype or tables_by_gemm" > gpurun_out/s13_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s13_tests.log; tail -15 gpurun_out/s13_tests.log
python - <<'PY'
import time, numpy as np, sys
sys.path.insert(0, '.')
import torch
from pygemma_b200 import lmm
n, m, c0 = 10000, 100000, 10
dev = 'cuda:0'
g = torch.Generator(device=dev); g.manual_seed(1)
mk = 2 * n
maf = torch.rand(mk, generator=g, device=dev) * 0.45 + 0.05
G = ((torch.rand(n, mk, generator=g, device=dev) < maf).double() + (torch.rand(n, mk, generator=g, device=dev) < maf).double())
sd = G.std(dim=0); sd[sd == 0] = 1; G = (G - G.mean(dim=0)) / sd
K = (G @ G.T / mk); K.diagonal().add_(1e-3); K = K.cpu().numpy(); del G
rng = np.random.default_rng(0)
W = np.c_[np.ones(n), rng.standard_normal((n, c0 - 1))]; y = rng.standard_normal(n)
X = rng.integers(0, 3, size=(n, m), dtype=np.int8)
for gp in (1, 2):
    for rep in range(2):
        t = time.time(); df = lmm.pygemma(y, X, W, K, gpus=gp); dt = time.time() - t
    print('gpus', gp, 'whole call incl. syevd %.3f s' % dt, {k: lmm.last_timing.get(k) for k in ('eig_ms', 'bcast_ms', 'design_ms', 'scan_wall_s')})
PY
