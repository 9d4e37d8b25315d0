# ncu --set full of the fused rotation kernel and of the REML-stage kernels (after the same commands ran clean without ncu)
timeout 300 python tools/prof_tc.py 10000 16384 > gpurun_out/prof_tc_plain.log 2>&1; echo "plain rc $?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:rotate_i8_tc2_kernel -c 1 --launch-skip 1 -o gpurun_out/tc2_r01c -f python tools/prof_tc.py 10000 16384 > gpurun_out/ncu_tc2.log 2>&1; echo "ncu rc $?"
python tools/ncu_summary.py gpurun_out/tc2_r01c.ncu-rep gpurun_out/ncu_r01_tc2_final_16384snps.json; cat gpurun_out/ncu_r01_tc2_final_16384snps.json
timeout 900 ncu --set full --clock-control none -k regex:"compress_dmma_kernel|fixed_xrow_kernel|reml_solve_kernel|pvalue_kernel" -c 4 --launch-skip 4 -o gpurun_out/reml_r01c -f python tools/prof_reml.py 10000 8192 10 > gpurun_out/ncu_reml.log 2>&1; echo "ncu rc $?"
python tools/ncu_summary.py gpurun_out/reml_r01c.ncu-rep gpurun_out/ncu_r01_reml_final_8192snps.json; cat gpurun_out/ncu_r01_reml_final_8192snps.json | head -120
