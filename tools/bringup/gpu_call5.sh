#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_ops_api_gpu.py -q -m gpu 2>&1 | tail -6
PERF_TAG=unified_singlepass timeout 600 python tools/attn_perf.py 2>&1 | tail -6
rm -f gpurun_out/ref_gpu.jsonl
timeout 1500 python tools/ref_gpu_paths.py > gpurun_out/ref_gpu.log 2>&1; tail -25 gpurun_out/ref_gpu.jsonl | cut -c1-400
