#!/bin/bash
# final round-1 evidence: full GPU suite, bench, stage timings, launch list, one ncu --set full of the band kernel
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -3 gpurun_out/bench_final.err; cut -c1-700 gpurun_out/bench_final.json
rm -f gpurun_out/stages.jsonl
timeout 900 python tools/stage_probe.py > gpurun_out/stages.log 2>&1; tail -2 gpurun_out/stages.log | cut -c1-300; cut -c1-200 gpurun_out/stages.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:attn_fwd|smse_|head_placement|plan_|permute|sort_|kmeans|dynmap|csq|sqnorm|km_|quant|absmax|qkv_prep|row_norm|qk_rope|glue" -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_ncu.log 2>&1
PROFILE_H=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 1 -c 1 -o gpurun_out/attn_band_v4 python tools/profile_attn.py > gpurun_out/ncu_band_v4.log 2>&1; tail -2 gpurun_out/ncu_band_v4.log
