#!/bin/bash
# where does the end-to-end time of ONE rank's shard of the 8-GPU strong-scaling problem go? (12 544 SNPs on one GPU)
mkdir -p gpurun_out
for snps in 12544 25088; do
timeout 600 python bench.py --snps $snps --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s24_$snps.json 2> gpurun_out/s24_$snps.err
python - $snps <<'PY'
import json,sys
j=[json.loads(l) for l in open(f"gpurun_out/s24_{sys.argv[1]}.json") if l.startswith("{")][-1]
e=j["e2e"]
print(sys.argv[1], "resident ms", round(j["ms_per_step"],3), "e2e ms", round(e["ms_per_step"],3), "pinned ms", round(j["e2e_pinned"]["ms_per_step"],3))
print("  last_call", {k:(round(v,4) if isinstance(v,float) else v) for k,v in e["last_call"].items() if k!="scan"})
print("  scan", e["last_call"]["scan"])
print("  pinned scan", j["e2e_pinned"]["timing_last_call"])
PY
done
