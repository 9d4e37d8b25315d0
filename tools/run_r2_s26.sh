#!/bin/bash
# A/B of the moment fusion on the default bench command (one B200): PG_FUSE_MOMENTS=0 (round-2 pipeline) vs fused
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "moment_fusion" --tb=short 2>&1 | tail -3
summ() { tail -1 $1 | python -c "
import json,sys
l=json.loads(sys.stdin.readline())
print(json.dumps({'value':l['value'],'ms':l['ms_per_step'],'e2e':l['e2e']['value'],'stages':l['roofline']['per_kernel_ms_last_step'],'eng':l['config']['rotation_engine'][:40],'spot':l['parity_spot'].get('max_rel'),'clk':l['clocks']['sm_mhz'],'pw':l['clocks'].get('power_w_max'),'e2e_ms': l['e2e'].get('ms_per_step')}))
"; }
for mode in ${MODES:-1 0 1 0}; do
  PG_FUSE_MOMENTS=$mode timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/fuse_ab_$mode.log 2>&1
  echo "c3 mode $mode rc $?"; summ gpurun_out/fuse_ab_$mode.log
done
for mode in ${MODES5:-1 0}; do
  PG_FUSE_MOMENTS=$mode timeout 300 python bench.py --config c5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/fuse_ab_c5_$mode.log 2>&1
  echo "c5 mode $mode rc $?"; summ gpurun_out/fuse_ab_c5_$mode.log
done
