#!/bin/bash
# final code: launch list of the timed region of the default bench command, and a fresh fuzz sweep (new seeds)
mkdir -p gpurun_out
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/s25_bench_pre_ncu.log 2>&1; echo "bench rc $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "pg_timed_resident/" --csv --log-file gpurun_out/launches_r02_final_bench_100000snps.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/s25_ncu_launch.log 2>&1; echo "ncu rc $?"
python tools/launch_summary.py gpurun_out/launches_r02_final_bench_100000snps.csv | tee gpurun_out/launches_r02_final_bench_100000snps.summary.txt
timeout 400 python tests/fuzz_parity.py 300 21 330 > gpurun_out/s25_fuzz21.log 2>&1; tail -1 gpurun_out/s25_fuzz21.log; grep "^FAIL" gpurun_out/s25_fuzz21.log | cut -c1-500 | head -5
timeout 400 python tests/fuzz_parity.py 300 22 330 > gpurun_out/s25_fuzz22.log 2>&1; tail -1 gpurun_out/s25_fuzz22.log; grep "^FAIL" gpurun_out/s25_fuzz22.log | cut -c1-500 | head -5
grep -h "flat_likelihood_rows" gpurun_out/s25_fuzz2*.log | cut -c1-300 | head
