#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tests/fuzz_parity.py 400 7 600 > gpurun_out/s15_fuzz.log 2>&1
tail -3 gpurun_out/s15_fuzz.log; grep -c "^FAIL" gpurun_out/s15_fuzz.log; grep "^FAIL" gpurun_out/s15_fuzz.log | head -12
