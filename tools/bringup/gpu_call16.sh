#!/bin/bash
cd "$(dirname "$0")/../.."
L=$PWD/sparse-videogen_b200/svgb200/_lib
for i in 1 2; do
SVGB200_LIB=$L/libsvgb200_base.so PERF_TAG=base timeout 300 python tools/ab_varblock.py | grep -E "aligned|QC1000|sample"
SVGB200_LIB=$L/libsvgb200_pp.so PERF_TAG=pingpong timeout 300 python tools/ab_varblock.py | grep -E "aligned|QC1000|sample"
PERF_TAG=split timeout 300 python tools/ab_varblock.py | grep -E "aligned|QC1000|sample"
done
