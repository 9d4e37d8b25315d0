#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_design.py 10000 10 > gpurun_out/s11_design_plain.log 2>&1; tail -2 gpurun_out/s11_design_plain.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s11_design_launches.csv python tools/prof_design.py 10000 10 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/s11_design_launches.csv | head -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/s11_bench_n2.json 2> gpurun_out/s11_bench_n2.err
python - <<PY
import json
j=json.loads([l for l in open('gpurun_out/s11_bench_n2.json') if l.startswith('{')][-1])
print('N=2', round(j['value']), round(j['ms_per_step'],3), 'e2e', round(j['e2e']['value']), round(j['e2e']['ms_per_step'],2), j['e2e']['last_call'])
PY
