"""One-kernel target for ncu: identify_dynamic_map at the HunyuanVideo shape (24 x 400 x 1000, D = 128)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "sparse-videogen_b200"))
from svgb200 import core  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
qc = torch.randn(24, 400, 128, device=dev, generator=g).bfloat16()
kc = torch.randn(24, 1000, 128, device=dev, generator=g).bfloat16()
ks = torch.randint(0, 300, (24, 1000), device=dev, generator=g, dtype=torch.int32)
for _ in range(3):
    m = core.dynamic_map(qc, kc, ks, 0.9, 100)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    m = core.dynamic_map(qc, kc, ks, 0.9, 100)
b.record()
torch.cuda.synchronize()
print("dynamic_map ms", a.elapsed_time(b) / 10, "kept", float(m.float().mean()))
