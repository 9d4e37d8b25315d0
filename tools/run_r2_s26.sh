#!/bin/bash
# A/B of the moment fusion on the default bench command (one B200): PG_FUSE_MOMENTS=0 (round-2 pipeline) vs fused
mkdir -p gpurun_out
timeout 300 python tools/try_fuse.py 1024,1500,5,1 2000,3000,10,4 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()[:300]); continue
    print(d['n'], d['engine_1'], 'fused_vs_unfused', max(d['fused_vs_unfused'].values()), 'vs_oracle', max(d['vs_oracle_1'].values()))
"
summ() { tail -1 $1 | python -c "
import json,sys
l=json.loads(sys.stdin.readline())
print(json.dumps({'value':l['value'],'ms':l['ms_per_step'],'e2e':l['e2e']['value'],'stages':l['roofline']['per_kernel_ms_last_step'],'eng':l['config']['rotation_engine'],'spot':l['parity_spot'].get('max_rel'),'clk':l['clocks']['sm_mhz'],'design_ms': l['setup']['design_tables_ms'], 'e2e_ms': l['e2e'].get('ms_per_step')}))
"; }
for mode in 1 0 1 0; do
  PG_FUSE_MOMENTS=$mode timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/fuse_ab_$mode.log 2>&1
  echo "c3 mode $mode rc $?"; summ gpurun_out/fuse_ab_$mode.log
done
for mode in 1 0; do
  PG_FUSE_MOMENTS=$mode timeout 300 python bench.py --config c5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/fuse_ab_c5_$mode.log 2>&1
  echo "c5 mode $mode rc $?"; summ gpurun_out/fuse_ab_c5_$mode.log
done
