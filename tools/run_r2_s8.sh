#!/bin/bash
# GPU session 8: the 4-CTA-cluster rotation (genotype tiles multicast to two CTA pairs): parity suite + bench A/B
mkdir -p gpurun_out
PG_TC_CLUSTER=4 timeout 600 python tools/prof_tc.py 10000 16384 > gpurun_out/s8_tc4_smoke.log 2>&1
echo "tc4 smoke rc=$?" >> gpurun_out/s8_tc4_smoke.log; tail -3 gpurun_out/s8_tc4_smoke.log
timeout 300 python tools/prof_tc.py 10000 16384 > gpurun_out/s8_tc2_smoke.log 2>&1; tail -1 gpurun_out/s8_tc2_smoke.log
PG_TC_CLUSTER=4 timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/s8_tests_tc4.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s8_tests_tc4.log; tail -8 gpurun_out/s8_tests_tc4.log
for cl in 4 2 4 2; do
PG_TC_CLUSTER=$cl timeout 600 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/s8_bench_cl$cl.json 2> gpurun_out/s8_bench_cl$cl.err
python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/s8_bench_cl$cl.json') if l.startswith('{')][-1])
    print('cluster $cl', round(j['value']), round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value']), j['roofline']['per_kernel_ms_last_step'], j['clocks'], j['parity_spot'].get('max_rel'))
except Exception as e:
    print('cluster $cl failed', e); print(open('gpurun_out/s8_bench_cl$cl.err').read()[-800:])
PY
done
