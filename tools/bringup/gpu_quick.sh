#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
timeout 600 python tools/attn_perf.py 2>&1 | grep -E "case" | cut -c1-150
