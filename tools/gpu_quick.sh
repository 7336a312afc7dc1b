#!/bin/bash
# quick iteration loop for attention-kernel changes: parity first, then the timing matrix
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py -q -m gpu -x 2>&1 | tail -8
timeout 600 python tools/attn_perf.py 2>&1 | tail -12
