timeout 300 python bench.py --steps 1 --warmup 3 > gpurun_out/bench_pre_ncu.log 2>&1; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_pre_ncu.log').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['per_kernel_ms_last_step'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "pg_timed_resident/" --csv --log-file gpurun_out/launches_r01_bench_100000snps.csv python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_launch.log 2>&1; echo "ncu rc $?"
python tools/launch_summary.py gpurun_out/launches_r01_bench_100000snps.csv | tee gpurun_out/launches_r01_bench_100000snps.summary.txt
