#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fp8_gpu.py tests/test_attention_gpu.py -q -m gpu -x -s 2>&1 | grep -E "passed|failed|fp8 attention|Error|error" | tail -12
timeout 600 python tools/attn_perf.py 2>&1 | grep -E "FP8|band|quantize|rho1.0" | tail -8
