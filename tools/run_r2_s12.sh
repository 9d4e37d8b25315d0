#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/s12_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s12_tests.log; tail -6 gpurun_out/s12_tests.log
timeout 300 python tools/prof_design.py 10000 10 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 > gpurun_out/s12_bench.json 2> gpurun_out/s12_bench.err
python - <<PY
import json
j=json.loads([l for l in open('gpurun_out/s12_bench.json') if l.startswith('{')][-1])
print('c3', round(j['value']), round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value']), round(j['e2e']['ms_per_step'],2), j['e2e']['last_call']['design_ms'], j['setup'], j['cpu_baseline']['value'], j['parity_spot'].get('max_rel'))
PY
timeout 600 python tools/multi_pheno_bench.py 10000 100000 10 1,4,16,64 > gpurun_out/s12_multi_pheno.jsonl 2>&1; tail -5 gpurun_out/s12_multi_pheno.jsonl
