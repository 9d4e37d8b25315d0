timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_fx.log 2>&1; echo "tests rc $?"
tail -3 gpurun_out/gpu_tests_fx.log
timeout 300 python tools/prof_reml.py 10000 50000 10
timeout 300 python tools/prof_reml.py 449 100000 6
timeout 300 python tools/prof_reml.py 10000 50000 40 grid
timeout 300 python bench.py > gpurun_out/bench_fx.log 2>&1; echo "bench rc $?"; tail -1 gpurun_out/bench_fx.log | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_fx.log').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['per_kernel_ms_last_step'])
PY
timeout 600 python tools/multi_pheno_bench.py 10000 100000 10 1,16 2>&1 | grep -v '^{"n"' | cut -c1-330
