#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/t_all.log; tail -8 gpurun_out/t_all.log
rm -f gpurun_out/stages.jsonl
timeout 900 python tools/stage_probe.py > gpurun_out/stages.log 2>&1; tail -40 gpurun_out/stages.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; tail -3 gpurun_out/bench2.err; cat gpurun_out/bench2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:attn_fwd|smse_|head_placement|plan_|permute|sort_|kmeans|dynmap|csq|sqnorm|km_" -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_ncu.log 2>&1
tail -2 gpurun_out/bench_ncu.log | cut -c1-300
