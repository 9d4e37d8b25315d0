#!/bin/bash
# round 2, GPU session 3 (8 GPUs): strong-scaling bench of the default config, then config c4 (n = 50 000 x 1 000 000 SNPs streamed from host)
mkdir -p gpurun_out
(nproc; free -g | head -2; nvidia-smi topo -m | head -12) > gpurun_out/s3_box.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/s3_bench_n8.json 2> gpurun_out/s3_bench_n8.err
echo "rc=$?" >> gpurun_out/s3_bench_n8.err
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --config c4 --steps 3 --warmup 3 > gpurun_out/s3_bench_c4_n8.json 2> gpurun_out/s3_bench_c4_n8.err
echo "rc=$?" >> gpurun_out/s3_bench_c4_n8.err
tail -c 400 gpurun_out/s3_bench_n8.err; tail -c 1200 gpurun_out/s3_bench_n8.json; tail -c 1500 gpurun_out/s3_bench_c4_n8.err; tail -c 1500 gpurun_out/s3_bench_c4_n8.json
