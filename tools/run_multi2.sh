set -x
timeout 900 python -m pytest tests -x -q -m gpu -k "trait or phenotype" > gpurun_out/gpu_tests_multi2.log 2>&1; echo "tests rc $?"
tail -5 gpurun_out/gpu_tests_multi2.log
timeout 600 python tools/multi_pheno_bench.py 10000 100000 10 16,16,32,64 > gpurun_out/multi_pheno2.log 2>&1; echo "mp rc $?"
grep -v '^{"n"' gpurun_out/multi_pheno2.log | cut -c1-400
timeout 500 python tools/fuzz_parity.py 200 2 300 > gpurun_out/fuzz2.log 2>&1; echo "fuzz rc $?"
grep "^FAIL" gpurun_out/fuzz2.log | cut -c1-300 | head; tail -1 gpurun_out/fuzz2.log
