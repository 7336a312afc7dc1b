#!/bin/bash
# under `gpurun --gpus N`: bench at N ranks with the head-parallel communication stream at high / default priority
cd "$(dirname "$0")/.."
N=$(nvidia-smi -L | wc -l)
for prio in -1 0 -1 0; do
  SVGB_HP_COMM_PRIORITY=$prio timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 20 --warmup 3 --no-extras > gpurun_out/bench_n${N}_prio$prio.json 2> gpurun_out/bench_n${N}_prio$prio.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n${N}_prio$prio.json').read().strip().splitlines()[-1])
print("prio $prio", d['value'], d['ms_per_step'], d['e2e']['value'])
PY
done
