#!/bin/bash
cd "$(dirname "$0")/../.."
L=$PWD/sparse-videogen_b200/svgb200/_lib
timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_fp8_gpu.py -x -q -m gpu 2>&1 | tail -2
for v in base tree d; do
SVGB200_LIB=$L/libsvgb200_$v.so PERF_TAG=$v timeout 300 python tools/ab_varblock.py | grep -E "uniform_QC400|ragged|aligned|dense|sample|band"
done
