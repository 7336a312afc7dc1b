#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_svg2_ops_gpu.py tests/test_fp8_gpu.py tests/test_fullsize_gpu.py tests/test_ops_api_gpu.py -x -q -m gpu > gpurun_out/pytest_pp.log 2>&1
echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_pp.log
for i in 1 2; do
SVGB200_LIB=$PWD/sparse-videogen_b200/svgb200/_lib/libsvgb200_base.so PERF_TAG=base timeout 300 python tools/ab_varblock.py
PERF_TAG=new timeout 300 python tools/ab_varblock.py
done
