set -x
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_multi3.log 2>&1; echo "tests rc $?"
tail -3 gpurun_out/gpu_tests_multi3.log
timeout 600 python tools/multi_pheno_bench.py 10000 50000 10 1,16,16,1,64 > gpurun_out/multi_pheno3.log 2>&1; echo "mp rc $?"
grep -v '^{"n"' gpurun_out/multi_pheno3.log | cut -c1-330
