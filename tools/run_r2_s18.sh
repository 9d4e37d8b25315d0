#!/bin/bash
# ncu: the solve kernel at the small-n shapes (c2: n=449 c0=6, c1: n=1940 c0=11)
mkdir -p gpurun_out
for cfg in "449 32768 6 c2" "1940 12226 11 c1"; do
  set -- $cfg
  timeout 300 python tools/prof_reml.py $1 $2 $3 > gpurun_out/s18_plain_$4.log 2>&1; echo "plain $4 rc $?"; tail -1 gpurun_out/s18_plain_$4.log | cut -c1-300
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:reml_solve_kernel -c 1 --launch-skip 1 -o gpurun_out/s18_solve_$4 -f python tools/prof_reml.py $1 $2 $3 > gpurun_out/s18_ncu_$4.log 2>&1; echo "ncu $4 rc $?"
  ncu -i gpurun_out/s18_solve_$4.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/s18_solve_$4_source.csv 2>/dev/null
  ncu -i gpurun_out/s18_solve_$4.ncu-rep --page raw --csv > gpurun_out/s18_solve_$4_raw.csv 2>/dev/null
  python tools/ncu_line_summary.py gpurun_out/s18_solve_$4_source.csv 40 > gpurun_out/s18_solve_$4_lines.txt; python tools/ncu_summary.py gpurun_out/s18_solve_$4.ncu-rep gpurun_out/s18_solve_$4.json; cat gpurun_out/s18_solve_$4.json
  head -50 gpurun_out/s18_solve_$4_lines.txt
done
rm -f gpurun_out/s18_solve_*_source.csv gpurun_out/s18_solve_*_raw.csv
ls -la gpurun_out/
