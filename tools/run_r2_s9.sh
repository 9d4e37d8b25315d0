#!/bin/bash
# GPU session 9: slab-in-shared-memory solver (PG_SOLVE_ZSM) A/B on c3 / c2 / c1 / c5 + parity suite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/s9_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s9_tests.log; tail -4 gpurun_out/s9_tests.log
for c in c2 c1 c3 c5; do for z in 1 0; do
PG_SOLVE_ZSM=$z timeout 600 python bench.py --config $c --steps 10 --no-cpu-baseline > gpurun_out/s9_bench_${c}_z$z.json 2> gpurun_out/s9_bench_${c}_z$z.err
python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/s9_bench_${c}_z$z.json') if l.startswith('{')][-1])
    print('$c zsm=$z', round(j['value']), round(j['ms_per_step'],3), 'e2e', round(j['e2e']['value']), j['roofline']['per_kernel_ms_last_step'], j['parity_spot'].get('max_rel'))
except Exception as e:
    print('$c zsm=$z failed', e); print(open('gpurun_out/s9_bench_${c}_z$z.err').read()[-800:])
PY
done; done
