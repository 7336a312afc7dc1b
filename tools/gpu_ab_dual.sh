#!/bin/bash
# A/B of the dual-item plan (SVGB_ATTN_PAIR) x softmax mapping (SVGB_ATTN_MAP) on the variable-block cases
cd "$(dirname "$0")/.."
for combo in "0 x" "0 0" "1 x" "1 0" "1 1"; do
  set -- $combo
  export SVGB_ATTN_PAIR=$1
  if [ "$2" = "x" ]; then unset SVGB_ATTN_MAP; else export SVGB_ATTN_MAP=$2; fi
  echo "== PAIR=$1 MAP=$2"
  PERF_TAG="pair$1_map$2" timeout 300 python tools/ab_varblock.py 2>&1 | grep -E "uniform_QC400|ragged|uniform_QC300|aligned_QC465|sample_mse|dense" 
done
unset SVGB_ATTN_PAIR SVGB_ATTN_MAP
echo "== tests PAIR=1 MAP=default"
timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_fullsize_gpu.py tests/test_ops_api_gpu.py tests/test_fp8_gpu.py -q -m gpu -x --no-header -p no:cacheprovider 2>&1 | tail -3
echo "== tests PAIR=1 MAP=0"
SVGB_ATTN_MAP=0 timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_fullsize_gpu.py tests/test_ops_api_gpu.py tests/test_fp8_gpu.py tests/test_configs_gpu.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -5
