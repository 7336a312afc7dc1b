#!/bin/bash
# A/B two library builds in one GPU call (same box, alternating) on the band case (+FP8)
for i in 1 2; do
  SVGB200_LIB=$PWD/sparse-videogen_b200/svgb200/_lib/libsvgb200_base.so PERF_TAG=base PERF_BAND_ONLY=1 python tools/attn_perf.py | grep case
  PERF_TAG=new PERF_BAND_ONLY=1 python tools/attn_perf.py | grep case
done
