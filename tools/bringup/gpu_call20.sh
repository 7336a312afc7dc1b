#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=$PWD/sparse-videogen_b200/svgb200/_lib
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_fp8_gpu.py -x -q -m gpu > gpurun_out/pytest_hyb.log 2>&1
echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_hyb.log
SVGB200_LIB=$L/libsvgb200_base.so PERF_TAG=base timeout 300 python tools/ab_varblock.py
PERF_TAG=splitTU timeout 300 python tools/ab_varblock.py
python - <<'PY'
import sys, json, torch
sys.path.insert(0, "sparse-videogen_b200"); sys.path.insert(0, ".")
from svgb200 import core
dev = torch.device("cuda:0")
H, V, D, KC = 24, 118800, 128, 1000
x = torch.randn(H, V, D, device=dev).bfloat16(); c = x[:, torch.randint(0, V, (KC,), device=dev)].contiguous()
xs = core.row_sqnorm(x)
for _ in range(2): core.kmeans_assign(x, c, xs)
torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): core.kmeans_assign(x, c, xs)
b.record(); torch.cuda.synchronize(); ms = a.elapsed_time(b) / 5
print(json.dumps({"case": "kmeans_assign_K1000", "ms": ms, "tflops": 2.0 * V * KC * D * H / ms / 1e9}))
PY
