#!/bin/bash
# multi-GPU session (gpurun --gpus N): head-parallel equality test + bench at N ranks
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 600 python -m pytest tests/test_configs_gpu.py -q -m gpu -k head_parallel --no-header -p no:cacheprovider 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 600 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
