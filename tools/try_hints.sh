#!/bin/bash
for hnt in 1 2 0; do
  echo "hints $hnt: $(PG_TC2_HINTS=$hnt python tools/prof_tc.py 10000 25088 2>&1 | tail -1 | grep -o "'rotate_ms': [0-9.]*")"
done
