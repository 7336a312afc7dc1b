#!/bin/bash
cd "$(dirname "$0")/.."
L=$PWD/sparse-videogen_b200/svgb200/_lib
echo "== sub-chunk pipeline, arrival-order issuer"
SVGB_ATTN_SUB=1 timeout 300 python -m pytest tests/test_attention_gpu.py tests/test_fullsize_gpu.py -q -m gpu -x --no-header -p no:cacheprovider 2>&1 | tail -3
for i in 1 2; do
  PERF_TAG=base PERF_BAND_ONLY=1 timeout 120 python tools/attn_perf.py | grep case | cut -c1-120
  SVGB_ATTN_SUB=1 PERF_TAG=sub PERF_BAND_ONLY=1 timeout 120 python tools/attn_perf.py | grep case | cut -c1-120
done
SVGB_ATTN_SUB=1 SVGB200_LIB=$L/libsvgb200_trace1.so TRACE_CASE=band TRACE_TAG=sub timeout 120 python tools/attn_trace.py 2>&1 | tail -4
bash tools/gpu_ab_dual.sh 2>&1 | grep -v "^\.\.\."
