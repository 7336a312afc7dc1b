"""k-means stage timings at HunyuanVideo-720p size (H=24, N=118800, D=128; K=1000 / 400).  JSON lines on stdout."""
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "sparse-videogen_b200"))
from svgb200 import core  # noqa: E402

dev = torch.device("cuda:0")


def t(fn, warm=2, iters=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


H, V, D = 24, 118800, 128
g = torch.Generator(device=dev).manual_seed(0)
cent = torch.randn(H, 1000, D, device=dev, generator=g) * 2
lab = torch.randint(0, 1000, (H, V), device=dev, generator=g)
x = (torch.gather(cent, 1, lab[:, :, None].expand(-1, -1, D)) + 0.5 * torch.randn(H, V, D, device=dev, generator=g)).bfloat16()
del cent, lab
xsq = core.row_sqnorm(x)
for K in (1000, 400):
    init = x[:, torch.randint(0, V, (K,), device=dev, generator=g)].contiguous()
    ms = t(lambda: core.kmeans_assign(x, init, xsq))
    print(json.dumps({"stage": f"assign_K{K}", "ms": ms, "tflops": 2.0 * V * K * D * H / ms / 1e9,
                      "lib": os.path.basename(os.environ.get("SVGB200_LIB", "default"))}), flush=True)
    if os.environ.get("KM_PROBE") == "assign":
        continue
    labels = core.kmeans_assign(x, init, xsq)
    ms = t(lambda: core.kmeans_update(x, labels, init))
    print(json.dumps({"stage": f"update_K{K} (incl. label sort)", "ms": ms, "gbs": x.numel() * 2 / ms / 1e6}), flush=True)
    ms = t(lambda: core.argsort_labels(labels, K))
    print(json.dumps({"stage": f"argsort_labels_K{K}", "ms": ms}), flush=True)
    ms = t(lambda: core.kmeans_run(x, init, 2))
    print(json.dumps({"stage": f"kmeans_run_K{K}_2it", "ms": ms}), flush=True)
if os.environ.get("KM_PROBE") == "assign":
    sys.exit(0)
ms = t(lambda: core.row_sqnorm(x))
print(json.dumps({"stage": "row_sqnorm (no longer on the Lloyd path)", "ms": ms}), flush=True)
qc = torch.randn(H, 400, D, device=dev, generator=g).bfloat16()
kc = torch.randn(H, 1000, D, device=dev, generator=g).bfloat16()
ks = torch.randint(1, 300, (H, 1000), device=dev, generator=g, dtype=torch.int32)
ms = t(lambda: core.dynamic_map(qc, kc, ks, 0.9, 100))
print(json.dumps({"stage": "dynamic_map 24x400x1000", "ms": ms}), flush=True)
