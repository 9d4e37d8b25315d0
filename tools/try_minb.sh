#!/bin/bash
# development probe: solve-kernel occupancy variants (PG_SOLVE_MINB) on the REML-only path
for b in 2 3 4; do
  echo "minb $b: $(PG_SOLVE_MINB=$b python tools/prof_reml.py ${1:-10000} ${2:-32768} ${3:-10} 2>&1 | tail -1 | grep -o "'reml_ms': [0-9.]*\|'compress_ms': [0-9.]*" | tr '\n' ' ')"
done
