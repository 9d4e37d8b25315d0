#!/bin/bash
# geometric block ramp for host-resident genotypes: e2e A/B at c3 (PG_RAMP=0 = no short blocks at all), then c1/c2/c5 e2e
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e=j["e2e"]; sc=e["last_call"]["scan"]
print(sys.argv[2], "resident", round(j["value"]), "e2e", round(e["value"]), "e2e ms", round(e["ms_per_step"],2), "pinned", round(j["e2e_pinned"]["value"]), "blocks", sc["n_blocks"], "scan total", sc["total_ms"], "rot", sc["rotate_ms"], "reml", sc["reml_ms"], "spot", j["parity_spot"]["max_rel"])
PY
}
for rep in 1 2; do
  PG_RAMP=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s20_c3_noramp_$rep.json 2>gpurun_out/s20.err; show gpurun_out/s20_c3_noramp_$rep.json "c3 PG_RAMP=0"
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s20_c3_ramp_$rep.json 2>gpurun_out/s20.err; show gpurun_out/s20_c3_ramp_$rep.json "c3 geometric"
done
for c in c2 c5; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s20_$c.json 2>gpurun_out/s20.err; show gpurun_out/s20_$c.json "$c geometric"
done
timeout 900 python -m pytest tests/test_gpu_fullsize.py -q -x -m gpu --tb=short 2>&1 | tail -3
