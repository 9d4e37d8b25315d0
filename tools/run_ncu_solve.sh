timeout 300 python tools/prof_reml.py 10000 8192 10 > gpurun_out/prof_reml_plain.log 2>&1; echo "plain rc $?"; tail -1 gpurun_out/prof_reml_plain.log | cut -c1-200
timeout 900 ncu --set full --import-source on --clock-control none -k regex:reml_solve_kernel -c 1 --launch-skip 1 -o gpurun_out/solve_r01b -f python tools/prof_reml.py 10000 8192 10 > gpurun_out/ncu_solve.log 2>&1; echo "ncu rc $?"
ncu -i gpurun_out/solve_r01b.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/solve_r01b_source.csv 2>/dev/null
python tools/ncu_line_summary.py gpurun_out/solve_r01b_source.csv 45
