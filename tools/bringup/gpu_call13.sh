#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/prep.jsonl
timeout 600 python -m pytest tests/test_prep_gpu.py -x -q -m gpu -k "glue or fused" > gpurun_out/pytest_prep2.log 2>&1
echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_prep2.log
timeout 600 python tools/prep_probe.py > gpurun_out/prep_probe.log 2>&1
echo "probe rc=$?"; cat gpurun_out/prep.jsonl
for i in 1 2; do
SVGB200_LIB=$PWD/sparse-videogen_b200/svgb200/_lib/libsvgb200_u8.so PREP_TAG=u8 PREP_FUSED_ONLY=1 timeout 300 python tools/prep_probe.py | grep fused_qkv
PREP_TAG=u4 PREP_FUSED_ONLY=1 timeout 300 python tools/prep_probe.py | grep fused_qkv
done
