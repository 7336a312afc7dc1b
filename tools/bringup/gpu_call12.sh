#!/bin/bash
# call 12: new prep / glue kernels (tests + probe), sample_mse merge change, bench for the e2e number
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/prep.jsonl
timeout 900 python -m pytest tests/test_prep_gpu.py tests/test_ops_api_gpu.py tests/test_svg2_ops_gpu.py -x -q -m gpu > gpurun_out/pytest_prep.log 2>&1
echo "pytest rc=$?" ; tail -5 gpurun_out/pytest_prep.log
timeout 600 python tools/prep_probe.py > gpurun_out/prep_probe.log 2>&1
echo "probe rc=$?"; cat gpurun_out/prep.jsonl
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_call12.json 2> gpurun_out/bench_call12.err
echo "bench rc=$?"; cat gpurun_out/bench_call12.json
