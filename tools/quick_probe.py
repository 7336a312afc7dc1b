"""GPU probe used during bring-up: times the attention kernel at HunyuanVideo-720p size under
(a) a synthetic 30%-density variable-block map and (b) the HY band mask, and a few layout ops.
Writes JSON lines to gpurun_out/probe.jsonl.  Not a bench (see bench.py)."""
import json
import math
import os
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "sparse-videogen_b200"))
sys.path.insert(0, str(ROOT))

from svgb200 import core  # noqa: E402

OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)
dev = torch.device("cuda:0")


def emit(**kw):
    print(json.dumps(kw), flush=True)
    with open(OUT / "probe.jsonl", "a") as f:
        f.write(json.dumps(kw) + "\n")


def time_fn(fn, warm=2, iters=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    H = int(os.environ.get("PROBE_H", 24))
    S, D = 119056, 128
    emit(what="device", name=torch.cuda.get_device_name(0), sm=core.device_check())
    g = torch.Generator(device="cpu").manual_seed(0)
    q = torch.randn(1, H, S, D, device=dev, dtype=torch.bfloat16)
    k = torch.randn(1, H, S, D, device=dev, dtype=torch.bfloat16)
    v = torch.randn(1, H, S, D, device=dev, dtype=torch.bfloat16)

    # (a) variable blocks: QC=400 / KC=1000 uniform-ish sizes, Bernoulli(0.3) map
    for QC, KC, rho in [(400, 1000, 0.3), (465, 931, 0.3), (1, 1, 1.0)]:
        def sizes(n):
            base = torch.full((H, n), S // n, dtype=torch.int32)
            base[:, : S - (S // n) * n] += 1
            return base
        row, col = sizes(QC), sizes(KC)
        bmap = torch.rand(H, QC, KC, generator=g) < rho
        bmap[:, :, 0] = True
        flops = 4.0 * D * (row.double()[:, :, None] * col.double()[:, None, :] * bmap).sum().item()
        if rho == 1.0 and H > 4:
            qq, kk, vv = q[:, :4], k[:, :4], v[:, :4]
            bm, rw, cl = bmap[:4], row[:4], col[:4]
            flops = flops * 4 / H
        else:
            qq, kk, vv, bm, rw, cl = q, k, v, bmap, row, col
        bm, rw, cl = bm.to(dev), rw.to(dev), cl.to(dev)
        t_plan, _ = time_fn(lambda: core.plan_varblock(bm, rw, cl, S))
        plan = core.plan_varblock(bm, rw, cl, S)
        t_med, t_min = time_fn(lambda: core.attn_fwd(qq, kk, vv, plan))
        emit(what="varblock", QC=QC, KC=KC, rho=rho, heads=qq.shape[1], ms=t_med, ms_min=t_min, plan_ms=t_plan,
             tflops=flops / t_med / 1e9, dense_equiv_tflops=4.0 * S * S * qq.shape[1] * D / t_med / 1e9)

    # (b) HY band mask at sparsity 0.30 -> W = floor(mul*P/128)*128
    ctx, F, P, plen = 256, 33, 3600, 60
    from math import floor, sqrt
    sp = 0.30
    seq = ctx + F * P
    s2 = (sp * seq * seq - 2 * seq * ctx) / (seq * seq)
    mul = seq * (1 - sqrt(1 - s2)) / P
    W = floor(mul * P / 128) * 128
    t0 = time.time()
    plan = core.plan_band(core.MASK_HY, F * P, F * P + plen, W, H, S, dev)
    torch.cuda.synchronize()
    t_plan = (time.time() - t0) * 1e3
    t_med, t_min = time_fn(lambda: core.attn_fwd(q, k, v, plan))
    # allowed pairs (analytic count on CPU, rows chunked)
    V, R = F * P, F * P + plen
    qi = torch.arange(S, dtype=torch.int64)
    lo = torch.clamp(qi - (W - 1), min=0)
    hi = torch.clamp(qi + (W - 1), max=V - 1)
    band = torch.clamp(hi - lo + 1, min=0)
    pairs = torch.where(qi < V, band + plen, torch.where(qi < R, torch.tensor(R), torch.tensor(S - R))).sum().item()
    flops = 4.0 * D * pairs * H
    emit(what="band_hy", W=W, mul=mul, ms=t_med, ms_min=t_min, plan_ms=t_plan, density=pairs / S / S,
         tflops=flops / t_med / 1e9)

    # (c) layout ops
    labels = torch.randint(0, 1000, (H, S), device=dev)
    t_sort, _ = time_fn(lambda: core.argsort_labels(labels, 1000))
    perm, _ = core.argsort_labels(labels, 1000)
    t_g, _ = time_fn(lambda: core.permute_gather(q, perm))
    t_s, _ = time_fn(lambda: core.permute_scatter(q, perm))
    gb = 2.0 * q.numel() * 2 / 1e9
    emit(what="layout", argsort_ms=t_sort, gather_ms=t_g, gather_gbs=gb / t_g * 1e3, scatter_ms=t_s,
         scatter_gbs=gb / t_s * 1e3)
    best = (torch.arange(H, device=dev) % 2).view(1, H)
    outs = [torch.empty_like(q) for _ in range(3)]
    t_p, _ = time_fn(lambda: core.head_placement([q, k, v], outs, best, ctx, F, P))
    emit(what="placement", ms=t_p, gbs=3 * gb / t_p * 1e3)


if __name__ == "__main__":
    main()
