#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 100 python -m pytest tests/test_attention_gpu.py -x -q -m gpu 2>&1 | tail -2
SVGB_ATTN_SUB=1 timeout 100 python -m pytest tests/test_attention_gpu.py -x -q -m gpu -k "band or dense or shd" 2>&1 | tail -3
SVGB_ATTN_SUB=1 PERF_TAG=sub PERF_BAND_ONLY=1 timeout 60 python tools/attn_perf.py | grep case | cut -c1-110
