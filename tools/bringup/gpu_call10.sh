#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
rm -f gpurun_out/stages.jsonl
timeout 900 python tools/stage_probe.py > gpurun_out/stages.log 2>&1; grep -E "svg2|svg1" gpurun_out/stages.jsonl | cut -c1-200
