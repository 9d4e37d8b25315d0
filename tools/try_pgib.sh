#!/bin/bash
# development probe: size of the int32 partial-product buffer (sub-block width of the int8 GEMM calls)
for g in 0.5 1 2 4 8; do
  echo "P_GIB $g: $(PG_P_GIB=$g python tools/prof_rot.py 10000 25088 10 2>&1 | tail -1 | grep -o "'rotate_ms': [0-9.]*\|'rotate_launches': [0-9]*" | tr '\n' ' ')"
done
