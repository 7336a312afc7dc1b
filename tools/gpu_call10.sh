#!/bin/bash
cd "$(dirname "$0")/.."
echo "== all gpu tests"
timeout 900 python -m pytest tests -q -m gpu -x --no-header -p no:cacheprovider 2>&1 | tail -6
echo "== perf: TAIL=1 vs TAIL=0"
for tl in 1 0; do
  SVGB_ATTN_TAIL=$tl PERF_TAG=tail$tl timeout 200 python tools/ab_varblock.py 2>&1 | grep -E "uniform_QC|ragged|aligned_QC465" | cut -c1-110
done
echo "== bench (no ref gpu)"
timeout 600 python bench.py --no-ref-gpu > gpurun_out/bench10.json 2> gpurun_out/bench10.err; tail -c 300 gpurun_out/bench10.err
