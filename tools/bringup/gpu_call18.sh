#!/bin/bash
cd "$(dirname "$0")/../.."
L=$PWD/sparse-videogen_b200/svgb200/_lib
SVGB200_LIB=$L/libsvgb200_trace.so timeout 200 python tools/attn_trace.py | tail -5
for i in 1 2; do
SVGB200_LIB=$L/libsvgb200_base.so PERF_TAG=base timeout 300 python tools/ab_varblock.py | grep -E "uniform_QC400|aligned|band"
PERF_TAG=stream timeout 300 python tools/ab_varblock.py | grep -E "uniform_QC400|aligned|band"
done
