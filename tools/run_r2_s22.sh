#!/bin/bash
# two GPUs, final code: the two-device tests (pg_multi_*, peer copy of U), the torchrun bench at N = 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi_device.py -q -m gpu --tb=short > gpurun_out/s22_multi_device.log 2>&1; echo "rc=$?" >> gpurun_out/s22_multi_device.log; tail -3 gpurun_out/s22_multi_device.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s22_bench_n2.json 2> gpurun_out/s22_bench_n2.err; echo "rc $?"
python - <<'PY'
import json
j=[json.loads(l) for l in open("gpurun_out/s22_bench_n2.json") if l.startswith("{")][-1]
print("N=2 value", round(j["value"]), "ms", j["ms_per_step"], "e2e", round(j["e2e"]["value"]), "check", j.get("multi_gpu_check"), "coll", j.get("collective_ms"), "spot", j["parity_spot"]["max_rel"])
PY
