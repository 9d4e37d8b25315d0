#!/bin/bash
# GPU session 10 (8 GPUs): strong-scaling bench with the final code, N = 8 and N = 4
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/s10_bench_n8.json 2> gpurun_out/s10_bench_n8.err
echo "rc=$?" >> gpurun_out/s10_bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/s10_bench_n4.json 2> gpurun_out/s10_bench_n4.err
echo "rc=$?" >> gpurun_out/s10_bench_n4.err
for n in 8 4; do python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/s10_bench_n$n.json') if l.startswith('{')][-1])
    print('N=$n', round(j['value']), round(j['ms_per_step'],3), 'e2e', round(j['e2e']['value']), round(j['e2e']['ms_per_step'],2), j['e2e']['last_call']['scan'].get('n_blocks'), j['collective_ms'], j['multi_gpu_check'], 'weak', round(j['weak']['value']), j['parity_spot'].get('max_rel'), j['parity_spot'].get('snps_checked'))
except Exception as e:
    print('N=$n failed', e); print(open('gpurun_out/s10_bench_n$n.err').read()[-1500:])
PY
done
