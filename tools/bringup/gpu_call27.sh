#!/bin/bash
cd "$(dirname "$0")/../.."
L=$PWD/sparse-videogen_b200/svgb200/_lib
timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_fp8_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu 2>&1 | tail -2
for i in 1 2 3; do
SVGB200_LIB=$L/libsvgb200_base.so PERF_TAG=base PERF_BAND_ONLY=1 timeout 120 python tools/attn_perf.py | grep case | cut -c1-120
PERF_TAG=earlyP PERF_BAND_ONLY=1 timeout 120 python tools/attn_perf.py | grep case | cut -c1-120
done
