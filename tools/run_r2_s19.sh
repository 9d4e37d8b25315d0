#!/bin/bash
# after the solver's padding skip (one-row last pass, node bound Kc): parity + c1 / c2 / c3 timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu --tb=short > gpurun_out/s19_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/s19_parity.log; tail -3 gpurun_out/s19_parity.log
for c in c1 c2 c3; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s19_bench_$c.json 2> gpurun_out/s19_bench_$c.err
  python - <<PY
import json
j=json.loads(open("gpurun_out/s19_bench_$c.json").read().strip().splitlines()[-1])
print("$c", round(j["value"]), j["ms_per_step"], j["roofline"].get("per_kernel_ms_last_step"), j["parity_spot"], j["e2e"]["value"])
PY
done
