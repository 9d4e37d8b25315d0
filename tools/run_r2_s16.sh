#!/bin/bash
# after the shared-memory opt-in fix: the c0 sweep test, the fuzz sweep again (same seed + a new one), full suite on both builds
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "every_covariate" --tb=short > gpurun_out/s16_c0sweep.log 2>&1
echo "c0 sweep rc=$?" >> gpurun_out/s16_c0sweep.log
timeout 700 python tests/fuzz_parity.py 400 7 500 > gpurun_out/s16_fuzz7.log 2>&1
timeout 500 python tests/fuzz_parity.py 300 11 350 > gpurun_out/s16_fuzz11.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/s16_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s16_tests.log
PYGEMMA_B200_LIB=$PWD/pygemma_b200/libpygemma_b200_dbg.so timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/s16_tests_dbg.log 2>&1
echo "pytest (PG_DEBUG_BOUNDS build) rc=$?" >> gpurun_out/s16_tests_dbg.log
tail -4 gpurun_out/s16_c0sweep.log; tail -1 gpurun_out/s16_fuzz7.log; grep "^FAIL" gpurun_out/s16_fuzz7.log | cut -c1-400 | head -5
tail -1 gpurun_out/s16_fuzz11.log; grep "^FAIL" gpurun_out/s16_fuzz11.log | cut -c1-400 | head -5
tail -4 gpurun_out/s16_tests.log; tail -3 gpurun_out/s16_tests_dbg.log
