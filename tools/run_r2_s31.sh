#!/bin/bash
# rotated-block stores of the unfused rotation: 16-byte (default), 32-byte (PG_TC2_HINTS bit 3)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "fused_tcgen05_rotation_is_bit_identical" 2>&1 | tail -2
PG_TC2_HINTS=10 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "fused_tcgen05_rotation_is_bit_identical or transposed_gemm" 2>&1 | tail -2
summ() { tail -1 $1 | python -c "
import json,sys
l=json.loads(sys.stdin.readline())
print(json.dumps({'value':l['value'],'ms':l['ms_per_step'],'stages':l['roofline']['per_kernel_ms_last_step'],'clk':l['clocks']['sm_mhz'],'spot':l['parity_spot'].get('max_rel')}))
"; }
for hints in 10 2 10 2 10 2; do
  PG_TC2_HINTS=$hints timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/hints_ab_$hints.log 2>&1
  echo "c3 hints $hints rc $?"; summ gpurun_out/hints_ab_$hints.log
done
