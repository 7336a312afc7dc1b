#!/bin/bash
cd "$(dirname "$0")/../.."
L=$PWD/sparse-videogen_b200/svgb200/_lib
for i in 1 2 3 4 5; do
SVGB200_LIB=$L/libsvgb200_base.so PERF_TAG=base PERF_BAND_ONLY=1 timeout 120 python tools/attn_perf.py | grep case | cut -c1-120
SVGB200_LIB=$L/libsvgb200_tree.so PERF_TAG=tree PERF_BAND_ONLY=1 timeout 120 python tools/attn_perf.py | grep case | cut -c1-120
done
