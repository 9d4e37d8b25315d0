#!/bin/bash
# N GPUs (argument), final code: torchrun bench
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s23_bench_n$N.json 2> gpurun_out/s23_bench_n$N.err; echo "rc $?"
python - $N <<'PY'
import json,sys
N=sys.argv[1]
j=[json.loads(l) for l in open(f"gpurun_out/s23_bench_n{N}.json") if l.startswith("{")][-1]
print("N=",N,"value", round(j["value"]), "ms", j["ms_per_step"], "e2e", round(j["e2e"]["value"]), j["e2e"]["ms_per_step"], "check", j.get("multi_gpu_check",{}).get("bit_identical_to_1gpu"), "gather ms", j.get("collective_ms",{}).get("gather_ms_per_step"), "spot", j["parity_spot"]["max_rel"], "weak", j.get("weak"))
PY
