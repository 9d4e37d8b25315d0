#!/bin/bash
# DRAM bytes of the rotation of one 25 088-SNP block (the bench's block size): planes pinned in the L2 set-aside vs single launch
mkdir -p gpurun_out
for v in 1 0; do
  PG_TC2_PERSIST=$v timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"rotate_i8_tc2_kernel" -c 60 --csv --log-file gpurun_out/s36_rot_$v.csv python tools/prof_tc.py 10000 25088 plain 10 > gpurun_out/s36_rot_$v.log 2>&1; echo "ncu persist=$v rc $?"
  python tools/ncu_metrics_sum.py gpurun_out/s36_rot_$v.csv rotate_i8_tc2_kernel 2 gpurun_out/ncu_r02_tc2_persist${v}_25088snps.json
done
