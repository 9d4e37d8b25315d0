#!/bin/bash
# round 2, GPU session 1: full GPU suite (incl. the full-size parity tests), the default bench line, L2 residency probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/s1_gpu.txt
free -g >> gpurun_out/s1_gpu.txt; nproc >> gpurun_out/s1_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --tb=short --durations=15 > gpurun_out/s1_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s1_tests.log
timeout 600 python bench.py > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
echo "bench rc=$?" >> gpurun_out/s1_bench.err
./tools/l2_probe > gpurun_out/s1_l2_probe.txt 2>&1
timeout 300 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s1_l2_ncu.csv ./tools/l2_probe > /dev/null 2>&1
tail -5 gpurun_out/s1_tests.log; tail -c 1500 gpurun_out/s1_bench.json
