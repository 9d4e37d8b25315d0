timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for cfg in "10000 50000 10" "449 100000 6" "1940 50000 11" "10000 50000 40 grid" "2000 50000 3"; do
  timeout 300 python tools/prof_reml.py $cfg 2>&1 | tail -1 | python -c "
import sys,ast
l=sys.stdin.read().strip()
d=ast.literal_eval(l[:l.rindex('}')+1])
print('$cfg: reml %.3f compress %.3f solve+p %.3f'%(d['reml_ms'], d['compress_ms'], d['reml_ms']-d['compress_ms']))"
done
timeout 300 python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_smem.log; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_smem.log").read())
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["per_kernel_ms_last_step"], d["clocks"]["sm_mhz"])
PY
