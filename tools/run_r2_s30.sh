#!/bin/bash
# automatic moment fusion (>= 24 linear columns): whole GPU suite on the new default, multi-trait passes with and without
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/s30_tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/s30_tests.log; tail -3 gpurun_out/s30_tests.log
for mode in 0 -1; do
  echo "== multi-trait, PG_FUSE_MOMENTS=$mode"
  PG_FUSE_MOMENTS=$mode timeout 400 python tools/multi_pheno_bench.py 10000 100000 10 8,16,32 2>&1 | grep '^{"q"' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in d.items() if k in ('q','total_ms','rotate_ms','reml_ms','compress_ms','rot_engine','n_nodes','tests_per_s','design_ms')})
"
done
