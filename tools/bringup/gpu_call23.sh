#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=$PWD/sparse-videogen_b200/svgb200/_lib
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_fp8_gpu.py tests/test_fullsize_gpu.py tests/test_svg2_ops_gpu.py tests/test_ops_api_gpu.py -x -q -m gpu > gpurun_out/pytest_comb.log 2>&1
echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_comb.log
SVGB200_LIB=$L/libsvgb200_head.so PERF_TAG=head timeout 300 python tools/ab_varblock.py
PERF_TAG=combined timeout 300 python tools/ab_varblock.py
