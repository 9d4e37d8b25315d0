# times the fused rotation with each pre-built library variant in variants/ (development helper)
cp pygemma_b200/libpygemma_b200.so /tmp/lib_orig.so
for rep in 1 2; do
for f in variants/*.so; do
  cp $f pygemma_b200/libpygemma_b200.so
  echo "== $f"
  timeout 300 python tools/prof_tc.py 10000 25088 2>&1 | tail -1 | python -c "
import sys,ast
l=sys.stdin.read().strip()
d=ast.literal_eval(l[:l.rindex('}')+1])
print('   rotate %.3f ms  %.1f Top/s'%(d['rotate_ms'], 14e8*25088/d['rotate_ms']/1e9))"
done
done
cp variants/lib_epi1.so pygemma_b200/libpygemma_b200.so
timeout 600 python -m pytest tests -x -q -m gpu -k "tcgen05 or float_dosages or bed or golden" 2>&1 | tail -2
cp /tmp/lib_orig.so pygemma_b200/libpygemma_b200.so
