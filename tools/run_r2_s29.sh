#!/bin/bash
# DRAM bytes and duration of the rotation kernel: unfused (plain / streaming stores of the rotated block), fused moments
mkdir -p gpurun_out
for v in "plain 2" "plain 6" "fuse 2"; do
  set -- $v
  PG_TC2_HINTS=$2 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"rotate_i8_tc2_kernel|moments_reduce_kernel|compress_dmma_kernel" -c 2 --launch-skip 2 --csv --log-file gpurun_out/s29_$1_$2.csv python tools/prof_tc.py 10000 16384 $1 10 > gpurun_out/s29_$1_$2.log 2>&1
  echo "== $v rc $?"; grep -v "^==" gpurun_out/s29_$1_$2.csv | python -c "
import csv,sys
for r in csv.reader(sys.stdin):
    if len(r) > 5 and r[0] != 'ID': print(r[4][:40], r[-3], r[-1])
"
done
