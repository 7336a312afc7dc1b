#!/bin/bash
cd "$(dirname "$0")/.."
L=$PWD/sparse-videogen_b200/svgb200/_lib
echo "== attention tests"
timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_fullsize_gpu.py tests/test_ops_api_gpu.py tests/test_reference_golden_gpu.py tests/test_svg2_ops_gpu.py -q -m gpu -x --no-header -p no:cacheprovider 2>&1 | tail -4
echo "== perf TAIL=1 vs 0"
for tl in 1 0; do
  SVGB_ATTN_TAIL=$tl PERF_TAG=tail$tl timeout 200 python tools/ab_varblock.py 2>&1 | grep -E "uniform_QC|ragged" | cut -c1-110
done
echo "== trace"
SVGB200_LIB=$L/libsvgb200_trace0.so TRACE_CASE=vb TRACE_TAG=tail2 timeout 200 python tools/attn_trace.py 2>&1 | tail -3
timeout 60 python tools/profile_dynmap.py
