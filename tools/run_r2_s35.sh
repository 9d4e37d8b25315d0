#!/bin/bash
# FINAL code (digit planes pinned in the L2 set-aside by default): GPU suite, DRAM bytes of the rotation, bench lines, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short > gpurun_out/gpu_tests_r02_final3.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/gpu_tests_r02_final3.log; tail -3 gpurun_out/gpu_tests_r02_final3.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"rotate_i8_tc2_kernel" -c 60 --csv --log-file gpurun_out/s35_rot.csv python tools/prof_tc.py 10000 16384 plain 10 > gpurun_out/s35_rot.log 2>&1; echo "ncu rc $?"
python tools/ncu_metrics_sum.py gpurun_out/s35_rot.csv rotate_i8_tc2_kernel 2 gpurun_out/ncu_r02_tc2_persist_16384snps.json
timeout 600 python bench.py > gpurun_out/bench_final3.log 2>&1; echo "bench rc $?"; tail -1 gpurun_out/bench_final3.log | cut -c1-250
timeout 300 python bench.py --config c5 --no-cpu-baseline > gpurun_out/bench_c5_final3.log 2>&1; echo "c5 rc $?"; tail -1 gpurun_out/bench_c5_final3.log | cut -c1-160
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "pg_timed_resident/" --csv --log-file gpurun_out/launches_r02_final3_bench_100000snps.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/s35_ncu_launch.log 2>&1; echo "ncu rc $?"
python tools/launch_summary.py gpurun_out/launches_r02_final3_bench_100000snps.csv | tee gpurun_out/launches_r02_final3_bench_100000snps.summary.txt
