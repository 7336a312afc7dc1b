#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=$PWD/sparse-videogen_b200/svgb200/_lib
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_fp8_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu > gpurun_out/pytest_lean.log 2>&1
echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_lean.log
for i in 1 2; do
SVGB200_LIB=$L/libsvgb200_base.so PERF_TAG=base timeout 300 python tools/ab_varblock.py | grep -E "uniform_QC400|aligned|band|ragged"
PERF_TAG=lean timeout 300 python tools/ab_varblock.py | grep -E "uniform_QC400|aligned|band|ragged"
done
