set -x
nvidia-smi -L | head -10
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err; echo "rc $?"
tail -1 gpurun_out/bench_n8.log
tail -5 gpurun_out/bench_n8.err
