#!/bin/bash
# GPU session 7: suite + bench after the two-power table-row skip, then the round-2 profiler evidence
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/s7_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s7_tests.log
tail -3 gpurun_out/s7_tests.log
for c in c3 c2 c1; do
timeout 600 python bench.py --config $c --steps 10 --no-cpu-baseline > gpurun_out/s7_bench_$c.json 2> gpurun_out/s7_bench_$c.err
python - <<PY
import json
j=json.loads([l for l in open('gpurun_out/s7_bench_$c.json') if l.startswith('{')][-1])
print('$c', round(j['value']), round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value']), j['roofline']['per_kernel_ms_last_step'], j['clocks'], j['parity_spot'].get('max_rel'))
PY
done
bash tools/run_r2_profiles.sh
