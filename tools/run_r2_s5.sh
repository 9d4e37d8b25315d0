#!/bin/bash
# round 2, GPU session 5: suite with the thread-per-SNP fixed phase; the same suite on the bounds-checked build; bench lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/s5_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s5_tests.log
PYGEMMA_B200_LIB=$PWD/pygemma_b200/libpygemma_b200_dbg.so timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/s5_tests_dbg.log 2>&1
echo "pytest (PG_DEBUG_BOUNDS build) rc=$?" >> gpurun_out/s5_tests_dbg.log
for c in c3 c5 c1 c2; do
  timeout 600 python bench.py --config $c --steps 5 --no-cpu-baseline > gpurun_out/s5_bench_$c.json 2> gpurun_out/s5_bench_$c.err
  echo "$c rc=$?" >> gpurun_out/s5_bench_$c.err
done
tail -6 gpurun_out/s5_tests.log; tail -4 gpurun_out/s5_tests_dbg.log
for c in c3 c5 c1 c2; do python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/s5_bench_$c.json') if l.startswith('{')][-1])
    print('$c', round(j['value']), round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value']), j['roofline']['per_kernel_ms_last_step'], j['parity_spot'].get('max_rel'))
except Exception as e:
    print('$c failed', e); print(open('gpurun_out/s5_bench_$c.err').read()[-1500:])
PY
done
