#!/bin/bash
# group size of the L2-pinned rotation vs the tail wave: 16 tiles (10.6 waves per launch at 49 SNP tiles), 18 (11.9), 21 (13.9)
mkdir -p gpurun_out
summ() { tail -1 $1 | python -c "
import json,sys
l=json.loads(sys.stdin.readline())
print(json.dumps({'value':l['value'],'ms':l['ms_per_step'],'stages':l['roofline']['per_kernel_ms_last_step'],'clk':l['clocks']['sm_mhz'],'spot':l['parity_spot'].get('max_rel')}))
"; }
for v in "16 40" "18 42" "21 48" "16 40" "18 42" "21 48"; do
  set -- $v
  PG_TC2_PERSIST_EG=$1 PG_TC2_PERSIST_MB=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/persist_eg_$1.log 2>&1
  echo "c3 group $1 setaside $2 MB rc $?"; summ gpurun_out/persist_eg_$1.log
done
