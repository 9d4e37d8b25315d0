#!/bin/bash
# round 2, GPU session 2 (2 GPUs): packing probe, 1-GPU bench (pageable e2e), 2-GPU strong-scaling bench
mkdir -p gpurun_out
(nproc; ./tools/pack_probe; echo no-stream; PG_NO_STREAM_COPY=1 ./tools/pack_probe; echo 8thr; PG_PACK_THREADS=8 ./tools/pack_probe; echo 32thr; PG_PACK_THREADS=32 ./tools/pack_probe) > gpurun_out/s2_pack.txt 2>&1
timeout 600 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/s2_bench_n1.json 2> gpurun_out/s2_bench_n1.err
echo "rc=$?" >> gpurun_out/s2_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/s2_bench_n2.json 2> gpurun_out/s2_bench_n2.err
echo "rc=$?" >> gpurun_out/s2_bench_n2.err
cat gpurun_out/s2_pack.txt; tail -c 600 gpurun_out/s2_bench_n1.err; tail -c 1500 gpurun_out/s2_bench_n2.err; tail -c 800 gpurun_out/s2_bench_n2.json
