#!/bin/bash
# eigen-major plane layout + (PG_TC2_PERSIST=1) one launch per eigen-tile group with its planes pinned in the L2 set-aside
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -m gpu -k "bit_identical or i8split or transposed or moment_fusion or c3_n10000_int8 or level_coded or packed_bed" --tb=short 2>&1 | tail -3
PG_TC2_PERSIST=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -m gpu -k "fused_tcgen05_rotation_is_bit_identical or c3_n10000_int8" --tb=short 2>&1 | tail -3
for v in 0 1; do
  PG_TC2_PERSIST=$v timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"rotate_i8_tc2_kernel" -c 20 --launch-skip ${SKIP:-0} --csv --log-file gpurun_out/s33_persist_$v.csv python tools/prof_tc.py 10000 16384 plain 10 > gpurun_out/s33_persist_$v.log 2>&1
  echo "== persist $v rc $?"; grep -v "^==" gpurun_out/s33_persist_$v.csv | python -c "
import csv,sys
tot={}
n=0
for r in csv.reader(sys.stdin):
    if len(r) > 5 and r[0] != 'ID':
        tot[r[-3]] = tot.get(r[-3], 0.0) + float(r[-1].replace(',', '')); n += 1
print(n // 3, 'launches (two scans)', {k: v / 2 for k, v in tot.items()})
"
done
summ() { tail -1 $1 | python -c "
import json,sys
l=json.loads(sys.stdin.readline())
print(json.dumps({'value':l['value'],'ms':l['ms_per_step'],'stages':l['roofline']['per_kernel_ms_last_step'],'clk':l['clocks']['sm_mhz'],'pw':l['clocks'].get('power_w_max'),'spot':l['parity_spot'].get('max_rel')}))
"; }
for p in 1 0 1 0; do
  PG_TC2_PERSIST=$p timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/persist_ab_$p.log 2>&1
  echo "c3 persist $p rc $?"; summ gpurun_out/persist_ab_$p.log
done
