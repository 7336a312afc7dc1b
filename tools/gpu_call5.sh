#!/bin/bash
cd "$(dirname "$0")/.."
L=$PWD/sparse-videogen_b200/svgb200/_lib
for rep in 1 2; do
for v in old poly25 default poly50; do
  if [ $v = default ]; then unset SVGB200_LIB; else export SVGB200_LIB=$L/libsvgb200_$v.so; fi
  PERF_TAG=$v PERF_BAND_ONLY=1 timeout 120 python tools/attn_perf.py | grep case | cut -c1-110
  [ $rep = 1 ] && PERF_TAG=$v timeout 200 python tools/ab_varblock.py 2>&1 | grep -E "uniform_QC400|ragged|aligned_QC465|dense|band_h12" | cut -c1-110
done; done
unset SVGB200_LIB
echo "== SUB=1 with default lib"
SVGB_ATTN_SUB=1 PERF_TAG=sub PERF_BAND_ONLY=1 timeout 120 python tools/attn_perf.py | grep case | cut -c1-110
echo "== tests (default lib)"
timeout 900 python -m pytest tests -q -m gpu -x --no-header -p no:cacheprovider 2>&1 | tail -4
