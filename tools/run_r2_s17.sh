#!/bin/bash
# timing experiment: does the fused rotation speed up when the genotype tiles cost fewer L2 -> SM bytes?
mkdir -p gpurun_out
export PYGEMMA_B200_LIB=$PWD/pygemma_b200/libpygemma_b200_exp.so
for rows in 0 64 32; do
  echo "== PG_TC2_EXP_A_ROWS=$rows" >> gpurun_out/s17_abytes.log
  PG_TC2_EXP_A_ROWS=$rows timeout 300 python tools/try_tc2_abytes.py 25088 >> gpurun_out/s17_abytes.log 2>&1
done
cat gpurun_out/s17_abytes.log
