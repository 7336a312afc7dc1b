#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_svg2_ops_gpu.py -q -m gpu 2>&1 | tail -40 > gpurun_out/t_svg2.log
tail -30 gpurun_out/t_svg2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; tail -3 gpurun_out/bench1.err; cat gpurun_out/bench1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:svgb -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_ncu.log 2>&1
tail -3 gpurun_out/bench_ncu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 1 -c 1 -o gpurun_out/attn_band python tools/profile_attn.py > gpurun_out/ncu_band.log 2>&1
tail -3 gpurun_out/ncu_band.log
ls -la gpurun_out
