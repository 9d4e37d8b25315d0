#!/bin/bash
# development probe: solve-kernel warps per CTA (L1 residency of the moments) on the REML-only path
for w in 8 6 5 4; do
  echo "warps $w: $(PG_SOLVE_WARPS=$w python tools/prof_reml.py ${1:-10000} ${2:-32768} ${3:-10} 2>&1 | tail -1 | grep -o "'reml_ms': [0-9.]*\|'compress_ms': [0-9.]*\|'n_nodes': [0-9]*" | tr '\n' ' ')"
done
