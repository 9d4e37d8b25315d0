#!/bin/bash
# GPU session 6: PG_OVERLAP experiment (solver of block b on 2 warps/SM under the rotation of block b+1) vs serial pipeline
mkdir -p gpurun_out
for mode in 0 1 0 1; do
  PG_OVERLAP=$mode timeout 600 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/s6_overlap$mode.json 2> gpurun_out/s6_overlap$mode.err
  python - <<PY
import json
j=json.loads([l for l in open('gpurun_out/s6_overlap$mode.json') if l.startswith('{')][-1])
print('PG_OVERLAP=$mode', round(j['value']), round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value']), j['roofline']['per_kernel_ms_last_step'], j['clocks'], j['parity_spot'].get('max_rel'))
PY
done
PG_OVERLAP=1 PG_OVERLAP_WARPS=1 timeout 600 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/s6_overlap1w1.json 2> gpurun_out/s6_overlap1w1.err
python - <<PY
import json
j=json.loads([l for l in open('gpurun_out/s6_overlap1w1.json') if l.startswith('{')][-1])
print('PG_OVERLAP=1 warps=1', round(j['value']), round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value']), j['roofline']['per_kernel_ms_last_step'], j['clocks'])
PY
