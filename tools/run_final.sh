python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.log 2>&1; echo "ref rc $?"; tail -1 gpurun_out/bench_ref_final.log | cut -c1-400
timeout 600 python bench.py > gpurun_out/bench_final.log 2>&1; echo "bench rc $?"; tail -1 gpurun_out/bench_final.log | cut -c1-250
