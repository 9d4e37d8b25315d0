#!/bin/bash
# GPU session 14: two SNPs per warp in the solver (PG_SOLVE_TILE=16 default vs 32): parity suite + A/B on c2 / c1 / c3
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/s14_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s14_tests.log; tail -6 gpurun_out/s14_tests.log
for c in c2 c1 c3; do for t in 16 32; do
PG_SOLVE_TILE=$t timeout 600 python bench.py --config $c --steps 10 --no-cpu-baseline > gpurun_out/s14_bench_${c}_t$t.json 2> gpurun_out/s14_bench_${c}_t$t.err
python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/s14_bench_${c}_t$t.json') if l.startswith('{')][-1])
    print('$c tile=$t', round(j['value']), round(j['ms_per_step'],3), 'e2e', round(j['e2e']['value']), j['roofline']['per_kernel_ms_last_step'], j['evals_per_snp'], j['parity_spot'].get('max_rel'))
except Exception as e:
    print('$c tile=$t failed', e); print(open('gpurun_out/s14_bench_${c}_t$t.err').read()[-800:])
PY
done; done
