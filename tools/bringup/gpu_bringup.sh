#!/bin/bash
# Bring-up sequence for one gpurun call: self-test first (isolates descriptor bugs), then parity
# tests, then the size probe.  Every step is bounded by `timeout` so a hung kernel cannot hold the box.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python -m pytest tests/test_attention_gpu.py -x -q -m gpu -k "selftest" 2>&1 | tail -40 > gpurun_out/t_selftest.log
cat gpurun_out/t_selftest.log | tail -15
timeout 900 python -m pytest tests/test_attention_gpu.py -q -m gpu -k "not selftest" 2>&1 | tail -60 > gpurun_out/t_attn.log
tail -25 gpurun_out/t_attn.log
timeout 600 python -m pytest tests/test_layout_gpu.py -q -m gpu 2>&1 | tail -40 > gpurun_out/t_layout.log
tail -15 gpurun_out/t_layout.log
rm -f gpurun_out/probe.jsonl
timeout 600 python tools/quick_probe.py > gpurun_out/probe.log 2>&1
tail -12 gpurun_out/probe.log
