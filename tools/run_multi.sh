set -x
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_multi.log 2>&1; echo "tests rc $?"
tail -5 gpurun_out/gpu_tests_multi.log
timeout 600 python tools/multi_pheno_bench.py 10000 100000 10 1,2,4,8,16 > gpurun_out/multi_pheno.log 2>&1; echo "mp rc $?"
tail -8 gpurun_out/multi_pheno.log
timeout 300 python bench.py > gpurun_out/bench_multi.log 2>&1; echo "bench rc $?"; tail -1 gpurun_out/bench_multi.log
