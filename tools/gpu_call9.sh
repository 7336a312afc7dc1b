#!/bin/bash
cd "$(dirname "$0")/.."
echo "== attention tests with the transposed tail kernel"
timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_fullsize_gpu.py tests/test_ops_api_gpu.py tests/test_reference_golden_gpu.py -q -m gpu -x --no-header -p no:cacheprovider 2>&1 | tail -15
echo "== perf: TAIL=1 vs TAIL=0"
for tl in 1 0; do
  SVGB_ATTN_TAIL=$tl PERF_TAG=tail$tl timeout 200 python tools/ab_varblock.py 2>&1 | grep -E "uniform_QC|ragged|aligned_QC465" | cut -c1-110
done
