#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/prep.jsonl
timeout 600 python -m pytest tests/test_prep_gpu.py tests/test_observability.py -x -q -m gpu > gpurun_out/pytest_prep3.log 2>&1
echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_prep3.log
timeout 600 python tools/prep_probe.py > gpurun_out/prep_probe.log 2>&1
echo "probe rc=$?"; grep wan gpurun_out/prep.jsonl
