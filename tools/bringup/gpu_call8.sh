#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -3 gpurun_out/bench_n$N.err | cut -c1-300; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['n_gpus','value','ms_per_step','gpu_launches']}, 'attn_ms',d['config']['attn_ms_per_call'],'roof',d['roofline']['achieved'],'e2e',d['e2e']['value'], d['clocks'])
PY
