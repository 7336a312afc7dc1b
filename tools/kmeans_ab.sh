#!/bin/bash
# A/B of assign-kernel build variants (sparse-videogen_b200/build.py --variant ...) inside ONE box session.
cd "$(dirname "$0")/.."
L=$PWD/sparse-videogen_b200/svgb200/_lib
timeout 600 python -m pytest tests/test_svg2_ops_gpu.py tests/test_reference_golden_gpu.py -q -m gpu -rf --no-header -p no:cacheprovider 2>&1 | tail -8
for v in ${KM_VARIANTS:-"" _kmOld "" _kmOld}; do
  [ "$v" = "-" ] && v=""
  SVGB200_LIB=$L/libsvgb200$v.so KM_PROBE=assign timeout 120 python tools/kmeans_probe.py 2>&1 | tail -2
done
