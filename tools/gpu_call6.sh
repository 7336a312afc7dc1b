#!/bin/bash
cd "$(dirname "$0")/.."
echo "== tests"
timeout 900 python -m pytest tests -q -m gpu -x --no-header -p no:cacheprovider 2>&1 | tail -4
echo "== arrival-order A/B (band + varblock)"
for rep in 1 2; do
  for o in 0 1; do
    SVGB_ATTN_ORDER=$o PERF_TAG=order$o PERF_BAND_ONLY=1 timeout 120 python tools/attn_perf.py | grep case | cut -c1-110
  done
done
SVGB_ATTN_ORDER=1 PERF_TAG=order1 timeout 200 python tools/ab_varblock.py 2>&1 | grep -E "uniform_QC400|ragged|aligned_QC465|dense|band_h12" | cut -c1-110
PERF_TAG=order0 timeout 200 python tools/ab_varblock.py 2>&1 | grep -E "uniform_QC400|ragged|aligned_QC465|dense|band_h12" | cut -c1-110
echo "== tests with ORDER=1"
SVGB_ATTN_ORDER=1 timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_fullsize_gpu.py tests/test_fp8_gpu.py -q -m gpu -x --no-header -p no:cacheprovider 2>&1 | tail -3
echo "== bench (no ref gpu)"
timeout 600 python bench.py --no-ref-gpu > gpurun_out/bench6.json 2> gpurun_out/bench6.err; tail -c 300 gpurun_out/bench6.err
