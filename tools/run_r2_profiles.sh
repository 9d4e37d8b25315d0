#!/bin/bash
# round 2 profiler evidence (run AFTER the same commands exited 0 without ncu):
#   1. ncu --set full of the fused rotation kernel (one launch, 16 384 SNPs at n = 10 000)
#   2. ncu --set full of the REML-stage kernels (compress, fixed x rows, fixed phase, solve, p-values)
#   3. launch list (gpu__time_duration.sum) of the timed region of the default bench command
mkdir -p gpurun_out
timeout 300 python tools/prof_tc.py 10000 16384 > gpurun_out/prof_tc_plain.log 2>&1; echo "plain tc rc $?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:rotate_i8_tc2_kernel -c 1 --launch-skip 1 -o gpurun_out/tc2_r02 -f python tools/prof_tc.py 10000 16384 > gpurun_out/ncu_tc2.log 2>&1; echo "ncu tc2 rc $?"
python tools/ncu_summary.py gpurun_out/tc2_r02.ncu-rep gpurun_out/ncu_r02_tc2_16384snps.json
timeout 300 python tools/prof_reml.py 10000 8192 10 > gpurun_out/prof_reml_plain.log 2>&1; echo "plain reml rc $?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"compress_dmma_kernel|fixed_xrow_kernel|fixed_phase_kernel|reml_solve_kernel|pvalue_kernel" -c 5 --launch-skip 5 -o gpurun_out/reml_r02 -f python tools/prof_reml.py 10000 8192 10 > gpurun_out/ncu_reml.log 2>&1; echo "ncu reml rc $?"
python tools/ncu_summary.py gpurun_out/reml_r02.ncu-rep gpurun_out/ncu_r02_reml_8192snps.json
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pre_ncu.log 2>&1; echo "bench rc $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "pg_timed_resident/" --csv --log-file gpurun_out/launches_r02_bench_100000snps.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launch list rc $?"
python tools/launch_summary.py gpurun_out/launches_r02_bench_100000snps.csv | tee gpurun_out/launches_r02_bench_100000snps.summary.txt
ls -la gpurun_out/*.ncu-rep
