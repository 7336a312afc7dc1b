"""A/B helper: variable-block attention cases dominated by single-tile items (bring-up tool)."""
import json
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
TAG = os.environ.get("PERF_TAG", "")


def t(fn, warm=2, iters=4):
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


H, S, D = 12, bench.S, bench.D
q, k, v = (torch.randn(1, H, S, D, device=dev, dtype=torch.bfloat16) for _ in range(3))
g = torch.Generator().manual_seed(0)


def uniform(n):
    b = torch.full((H, n), S // n, dtype=torch.int32)
    b[:, : S - (S // n) * n] += 1
    return b


def ragged(n):
    out = []
    for _ in range(H):
        cuts = torch.sort(torch.randperm(S - 1, generator=g)[: n - 1] + 1)[0]
        out.append(torch.diff(torch.cat([torch.tensor([0]), cuts, torch.tensor([S])])).to(torch.int32))
    return torch.stack(out)


for name, QC, KC, rho, fn in [("uniform", 400, 1000, 0.3, uniform), ("ragged", 400, 1000, 0.3, ragged),
                              ("uniform", 1000, 1000, 0.3, uniform), ("uniform", 300, 1000, 0.2, uniform)]:
    row, col = fn(QC), fn(KC)
    bm = torch.rand(H, QC, KC, generator=g) < rho
    bm[:, :, 0] = True
    fl = 4.0 * D * (row.double()[:, :, None] * col.double()[:, None, :] * bm).sum().item()
    pl = core.plan_varblock(bm.to(dev), row.to(dev), col.to(dev), S)
    ms = t(lambda: core.attn_fwd(q, k, v, pl))
    print(json.dumps(dict(tag=TAG, case=f"{name}_QC{QC}_KC{KC}_rho{rho}", ms=round(ms, 3), tflops=round(fl / ms / 1e9, 1))), flush=True)
# full single tiles, full 128-column chunks (plain softmax path): S2 = 128 * 930
S2 = 128 * 930
q2, k2, v2 = (x[:, :, :S2].contiguous() for x in (q, k, v))
for QC2, KC2 in [(930, 930), (465, 930)]:
    row = torch.full((H, QC2), S2 // QC2, dtype=torch.int32)
    col = torch.full((H, KC2), S2 // KC2, dtype=torch.int32)
    bm = torch.rand(H, QC2, KC2, generator=g) < 0.3
    bm[:, :, 0] = True
    fl = 4.0 * D * (row.double()[:, :, None] * col.double()[:, None, :] * bm).sum().item()
    pl = core.plan_varblock(bm.to(dev), row.to(dev), col.to(dev), S2)
    ms = t(lambda: core.attn_fwd(q2, k2, v2, pl))
    print(json.dumps(dict(tag=TAG, case=f"aligned_QC{QC2}_KC{KC2}", ms=round(ms, 3), tflops=round(fl / ms / 1e9, 1))), flush=True)
del q2, k2, v2
one = torch.full((4, 1), S, dtype=torch.int32, device=dev)
pl = core.plan_varblock(torch.ones(4, 1, 1, dtype=torch.bool, device=dev), one, one, S)
qd, kd, vd = (x[:, :4].contiguous() for x in (q, k, v))
ms = t(lambda: core.attn_fwd(qd, kd, vd, pl), warm=1, iters=2)
print(json.dumps(dict(tag=TAG, case="dense_h4", ms=round(ms, 3), tflops=round(4.0 * D * S * S * 4 / ms / 1e9, 1))), flush=True)
del qd, kd, vd
rows = torch.randint(0, 10000, (64,), device=dev, dtype=torch.int32)
ms = t(lambda: core.sample_mse(q.view(H, S, D), k.view(H, S, D), v.view(H, S, D), rows, 0, bench.CTX, bench.F, bench.P))
print(json.dumps(dict(tag=TAG, case="sample_mse_h12", ms=round(ms, 3))), flush=True)
W, _ = bench.band_width()
plan = core.plan_band(core.MASK_HY, bench.F * bench.P, bench.F * bench.P + bench.PROMPT_LEN, W, H, S, dev)
ms = t(lambda: core.attn_fwd(q, k, v, plan))
print(json.dumps(dict(tag=TAG, case="band_h12", ms=round(ms, 3), tflops=round(4.0 * D * bench.band_pairs(W) * H / ms / 1e9, 1))), flush=True)
