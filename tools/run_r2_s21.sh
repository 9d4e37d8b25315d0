#!/bin/bash
# final validation of the round's code on one GPU: smoke, full GPU suite on both builds, reference arm, default bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s21_smoke.log 2>&1; tail -2 gpurun_out/s21_smoke.log
timeout 1500 python -m pytest tests -q -m gpu --tb=short > gpurun_out/s21_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s21_tests.log; tail -3 gpurun_out/s21_tests.log
PYGEMMA_B200_LIB=$PWD/pygemma_b200/libpygemma_b200_dbg.so timeout 1500 python -m pytest tests -q -m gpu --tb=short > gpurun_out/s21_tests_dbg.log 2>&1; echo "pytest (PG_DEBUG_BOUNDS build) rc=$?" >> gpurun_out/s21_tests_dbg.log; tail -3 gpurun_out/s21_tests_dbg.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s21_bench_ref.json 2> gpurun_out/s21_bench_ref.err; echo "ref rc $?"; tail -1 gpurun_out/s21_bench_ref.json | cut -c1-500
timeout 900 python bench.py > gpurun_out/s21_bench.json 2> gpurun_out/s21_bench.err; echo "bench rc $?"; tail -1 gpurun_out/s21_bench.json | cut -c1-400
for c in c1 c2 c5; do timeout 600 python bench.py --config $c --no-cpu-baseline > gpurun_out/s21_bench_$c.json 2>gpurun_out/s21_bench_$c.err; tail -1 gpurun_out/s21_bench_$c.json | cut -c1-200; done
