#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json | cut -c1-2500
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -5 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json | cut -c1-2500
