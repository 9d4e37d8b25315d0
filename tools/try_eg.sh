#!/bin/bash
# development probe: eigen-tile group size of the fused rotation kernel (L2 residency)
for g in 6 12 24 48 313; do
  echo "EG $g: $(PG_TC2_EG=$g python tools/prof_tc.py 10000 25088 2>&1 | tail -1 | grep -o "'rotate_ms': [0-9.]*")"
done
