#!/bin/bash
# moment fusion: bounds-checked build, fuzz sweep with fusion on, ncu --set full of the fused kernel (DRAM traffic)
mkdir -p gpurun_out
DBG=$PWD/pygemma_b200/libpygemma_b200_dbg.so
PYGEMMA_B200_LIB=$DBG timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "moment_fusion" --tb=short > gpurun_out/s28_fusion_dbg.log 2>&1
echo "fusion test (PG_DEBUG_BOUNDS build) rc=$?" | tee -a gpurun_out/s28_fusion_dbg.log; tail -2 gpurun_out/s28_fusion_dbg.log
PG_FUSE_MOMENTS=1 timeout 400 python tests/fuzz_parity.py 400 31 330 > gpurun_out/s28_fuzz31_fused.log 2>&1; tail -1 gpurun_out/s28_fuzz31_fused.log; grep "^FAIL" gpurun_out/s28_fuzz31_fused.log | cut -c1-500 | head -5
PG_FUSE_MOMENTS=1 PYGEMMA_B200_LIB=$DBG timeout 300 python tests/fuzz_parity.py 150 32 200 > gpurun_out/s28_fuzz32_fused_dbg.log 2>&1; tail -1 gpurun_out/s28_fuzz32_fused_dbg.log; grep "^FAIL" gpurun_out/s28_fuzz32_fused_dbg.log | cut -c1-500 | head -5
timeout 300 python tools/prof_tc.py 10000 16384 fuse 10 > gpurun_out/s28_prof_plain.log 2>&1; echo "plain rc $?"; tail -1 gpurun_out/s28_prof_plain.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"rotate_i8_tc2_kernel|moments_reduce_kernel" -c 2 --launch-skip 2 -o gpurun_out/tc2_fused_r02 -f python tools/prof_tc.py 10000 16384 fuse 10 > gpurun_out/s28_ncu.log 2>&1; echo "ncu rc $?"
python tools/ncu_summary.py gpurun_out/tc2_fused_r02.ncu-rep gpurun_out/ncu_r02_tc2_fused_16384snps.json; cat gpurun_out/ncu_r02_tc2_fused_16384snps.json | head -80
