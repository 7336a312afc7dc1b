"""Small driver for `ncu --set full`: a few launches of the attention kernel at HunyuanVideo-720p
sequence length with 2 heads (band mask rho=0.30, or variable blocks with PROFILE_MODE=varblock)."""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "sparse-videogen_b200"))
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from svgb200 import core  # noqa: E402

dev = torch.device("cuda:0")
H = int(os.environ.get("PROFILE_H", 2))
S, D = bench.S, bench.D
q, k, v = (torch.randn(1, H, S, D, device=dev, dtype=torch.bfloat16) for _ in range(3))
if os.environ.get("PROFILE_MODE", "band") == "band":
    W, _ = bench.band_width()
    plan = core.plan_band(core.MASK_HY, bench.F * bench.P, bench.F * bench.P + bench.PROMPT_LEN, W, H, S, dev)
else:
    QC, KC = (int(x) for x in os.environ.get("PROFILE_QCKC", "465,931").split(","))
    g = torch.Generator().manual_seed(0)

    def sizes(n):
        b = torch.full((H, n), S // n, dtype=torch.int32)
        b[:, : S - (S // n) * n] += 1
        return b
    bm = torch.rand(H, QC, KC, generator=g) < 0.3
    plan = core.plan_varblock(bm.to(dev), sizes(QC).to(dev), sizes(KC).to(dev), S)
for _ in range(int(os.environ.get("PROFILE_ITERS", 3))):
    o = core.attn_fwd(q, k, v, plan)
torch.cuda.synchronize()
print("done", float(o.float().abs().mean()))
