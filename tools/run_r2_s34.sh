#!/bin/bash
# L2 persistence of the plane group: group size / set-aside sweep on the default bench command
mkdir -p gpurun_out
summ() { tail -1 $1 | python -c "
import json,sys
l=json.loads(sys.stdin.readline())
print(json.dumps({'value':l['value'],'ms':l['ms_per_step'],'stages':l['roofline']['per_kernel_ms_last_step'],'clk':l['clocks']['sm_mhz'],'pw':l['clocks'].get('power_w_max'),'e2e':l['e2e']['value']}))
"; }
for v in "0 24 64" "1 16 40" "1 20 48" "1 24 56" "0 24 64" "1 16 40" "1 20 48" "1 24 56"; do
  set -- $v
  PG_TC2_PERSIST=$1 PG_TC2_EG=$2 PG_TC2_PERSIST_MB=$3 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/persist_ab_$1_$2.log 2>&1
  echo "c3 persist $1 group $2 setaside $3 MB rc $?"; summ gpurun_out/persist_ab_$1_$2.log
done
