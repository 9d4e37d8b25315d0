#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short --durations=8 > gpurun_out/s4_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s4_tests.log
timeout 600 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/s4_bench_n1.json 2> gpurun_out/s4_bench_n1.err
tail -25 gpurun_out/s4_tests.log; tail -c 300 gpurun_out/s4_bench_n1.err
