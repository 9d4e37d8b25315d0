#!/bin/bash
# FINAL code of round 2 (session 2): smoke, whole GPU suite, reference arm, default bench, per-config lines, launch list
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python -m pytest tests -q -m gpu --tb=short > gpurun_out/gpu_tests_r02_final2.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/gpu_tests_r02_final2.log; tail -3 gpurun_out/gpu_tests_r02_final2.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final2.log 2>&1; echo "ref rc $?"; tail -1 gpurun_out/bench_ref_final2.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench_final2.log 2>&1; echo "bench rc $?"; tail -1 gpurun_out/bench_final2.log | cut -c1-250
for cfg in c1 c2 c5; do
  timeout 300 python bench.py --config $cfg --no-cpu-baseline > gpurun_out/bench_${cfg}_final2.log 2>&1; echo "$cfg rc $?"; tail -1 gpurun_out/bench_${cfg}_final2.log | cut -c1-160
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "pg_timed_resident/" --csv --log-file gpurun_out/launches_r02_final2_bench_100000snps.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/s32_ncu_launch.log 2>&1; echo "ncu rc $?"
python tools/launch_summary.py gpurun_out/launches_r02_final2_bench_100000snps.csv | tee gpurun_out/launches_r02_final2_bench_100000snps.summary.txt
