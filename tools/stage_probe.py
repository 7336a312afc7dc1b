"""Per-stage CUDA-event timings at HunyuanVideo-720p size (SVG1 and SVG2 paths).  JSON lines to
gpurun_out/stages.jsonl.  Bring-up tool, not the bench."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "sparse-videogen_b200"))
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from svgb200 import core  # noqa: E402
from svgb200.models import hyvideo as hy  # noqa: E402

dev = torch.device("cuda:0")
OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)


def emit(**kw):
    print(json.dumps(kw), flush=True)
    with open(OUT / "stages.jsonl", "a") as f:
        f.write(json.dumps(kw) + "\n")


def t(fn, warm=1, iters=3):
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


H, S, D, F, P, CTX = 24, bench.S, bench.D, bench.F, bench.P, bench.CTX
q, k, v = (torch.randn(1, H, S, D, device=dev, dtype=torch.bfloat16) for _ in range(3))
gb = q.numel() * 2 / 1e9

# ---- SVG1 stages
svg1 = hy.HunyuanSVG1Core(CTX, bench.PROMPT_LEN, F, P, H, D, bench.SPARSITY, dev)
rows = torch.randint(0, 10000, (64,))
emit(stage="svg1.sample_mse", ms=t(lambda: svg1.sample_mse(q, k, v, rows)))
best = (torch.arange(H, device=dev) % 2).view(1, H)
outs = [torch.empty_like(q) for _ in range(3)]
ms = t(lambda: core.head_placement([q, k, v], outs, best, CTX, F, P))
emit(stage="svg1.placement_qkv", ms=ms, gbs=6 * gb / ms * 1e3)
ms = t(lambda: core.head_placement([q], outs[:1], best, CTX, F, P, inverse=True))
emit(stage="svg1.inverse_placement", ms=ms, gbs=2 * gb / ms * 1e3)
emit(stage="svg1.band_attention", ms=t(lambda: core.attn_fwd(q, k, v, svg1.block_mask.plan)))
emit(stage="svg1.sparse_core_total", ms=t(lambda: svg1.sparse_core(q, k, v, rows)))
del outs

# ---- SVG2 stages (QC=400, KC=1000, clustered data so k-means has structure)
V = F * P
QC, KC = 400, 1000
g = torch.Generator(device=dev).manual_seed(0)
cent = torch.randn(H, KC, D, device=dev, generator=g) * 2
lab = torch.randint(0, KC, (H, V), device=dev, generator=g)
xk = (torch.gather(cent, 1, lab[:, :, None].expand(-1, -1, D))
      + 0.5 * torch.randn(H, V, D, device=dev, generator=g)).bfloat16()
del cent, lab
xq = xk.clone()
init_k = xk[:, torch.randint(0, V, (KC,), device=dev, generator=g)].contiguous()
init_q = xq[:, torch.randint(0, V, (QC,), device=dev, generator=g)].contiguous()
xsq = core.row_sqnorm(xk)
emit(stage="svg2.row_sqnorm", ms=t(lambda: core.row_sqnorm(xk)))
ms = t(lambda: core.kmeans_assign(xk, init_k, xsq))
emit(stage="svg2.assign_K1000", ms=ms, tflops=2.0 * V * KC * D * H / ms / 1e9)
ms = t(lambda: core.kmeans_assign(xq, init_q, xsq))
emit(stage="svg2.assign_K400", ms=ms, tflops=2.0 * V * QC * D * H / ms / 1e9)
labels = core.kmeans_assign(xk, init_k, xsq)
ms = t(lambda: core.kmeans_update(xk, labels, init_k))
emit(stage="svg2.update_K1000", ms=ms, gbs=gb * V / S / ms * 1e3)
emit(stage="svg2.kmeans_run_K1000_2it", ms=t(lambda: core.kmeans_run(xk, init_k, 2)))
emit(stage="svg2.kmeans_run_K400_2it", ms=t(lambda: core.kmeans_run(xq, init_q, 2)))
emit(stage="svg2.kmeans_run_K1000_50it(tol)", ms=t(lambda: core.kmeans_run(xk, init_k, 50), warm=0, iters=1))
ql, qc, qs, _ = core.kmeans_run(xq, init_q, 2)
kl, kc, ks, _ = core.kmeans_run(xk, init_k, 2)
emit(stage="svg2.dynamic_map", ms=t(lambda: core.dynamic_map(qc, kc, ks, 0.9, 100)))
emit(stage="svg2.argsort_labels", ms=t(lambda: core.argsort_labels(kl, KC)))
perm, _ = core.argsort_labels(kl, KC)
full_perm = torch.cat([perm, torch.arange(V, S, device=dev, dtype=torch.int32).expand(H, CTX)], 1)
ms = t(lambda: core.permute_gather(k, full_perm))
emit(stage="svg2.permute_gather", ms=ms, gbs=2 * gb / ms * 1e3)
sap = hy.HunyuanSAPCore(CTX, F, P, num_q_centroids=QC, num_k_centroids=KC, top_p_kmeans=0.9, min_kc_ratio=0.1,
                        kmeans_iter_init=50, kmeans_iter_step=2, prompt_length=bench.PROMPT_LEN)
qq = torch.cat([xq, q[0, :, V:]], 1)[None].contiguous()
kk = torch.cat([xk, k[0, :, V:]], 1)[None].contiguous()
del xq, xk, q, k
emit(stage="svg2.sap_core_first_call(50it)", ms=t(lambda: sap.sparse_core(qq, kk, v), warm=0, iters=1))
emit(stage="svg2.sap_core_step_call(2it)", ms=t(lambda: sap.sparse_core(qq, kk, v), warm=0, iters=2))
dens = core.density(sap.last["dynamic_map"], sap.last["q_sizes"], sap.last["k_sizes"])
emit(stage="svg2.density", mean=float(dens.mean()), min=float(dens.min()), max=float(dens.max()))
plan = core.plan_varblock(sap.last["dynamic_map"], sap.last["q_sizes"], sap.last["k_sizes"], S)
qp = core.permute_gather(qq, sap.last["q_sorted_indices"])
kp = core.permute_gather(kk, sap.last["k_sorted_indices"])
vp = core.permute_gather(v, sap.last["k_sorted_indices"])
ms = t(lambda: core.attn_fwd(qp, kp, vp, plan, o_rows=sap.last["q_sorted_indices"]))
fl = 4.0 * D * (sap.last["q_sizes"].double()[:, :, None] * sap.last["k_sizes"].double()[:, None, :]
                * sap.last["dynamic_map"]).sum().item()
emit(stage="svg2.attention_on_kmeans_map", ms=ms, tflops=fl / ms / 1e9, density=fl / (4.0 * D * H * S * S))

# ---- Wan 2.1 720p shape (BASELINE config 2: H=40, S=75600 = 21 x 3600, no text), SVG2: QC=300 / KC=1000
del qq, kk, v, qp, kp, vp, plan, sap
torch.cuda.empty_cache()
from svgb200.models import wan  # noqa: E402

Hw, Fw, Pw = 40, 21, 3600
Sw = Fw * Pw
gw = torch.Generator(device=dev).manual_seed(1)
centw = torch.randn(Hw, 1000, D, device=dev, generator=gw) * 2
labw = torch.randint(0, 1000, (Hw, Sw), device=dev, generator=gw)
kw = (torch.gather(centw, 1, labw[:, :, None].expand(-1, -1, D)) + 0.5 * torch.randn(Hw, Sw, D, device=dev, generator=gw)).bfloat16()[None]
qw = kw.clone()
vw = torch.randn(1, Hw, Sw, D, device=dev, generator=gw).bfloat16()
del centw, labw
sapw = wan.WanSAPCore(Fw, Pw, num_q_centroids=300, num_k_centroids=1000, top_p_kmeans=0.9, min_kc_ratio=0.1,
                      kmeans_iter_init=50, kmeans_iter_step=2)
emit(stage="wan.sap_core_first_call(50it)", ms=t(lambda: sapw.sparse_core(qw, kw, vw), warm=0, iters=1))
emit(stage="wan.sap_core_step_call(2it)", ms=t(lambda: sapw.sparse_core(qw, kw, vw), warm=0, iters=2))
dw = core.density(sapw.last["dynamic_map"], sapw.last["q_sizes"], sapw.last["k_sizes"])
planw = core.plan_varblock(sapw.last["dynamic_map"], sapw.last["q_sizes"], sapw.last["k_sizes"], Sw)
qpw = core.permute_gather(qw, sapw.last["q_sorted_indices"])
kpw = core.permute_gather(kw, sapw.last["k_sorted_indices"])
vpw = core.permute_gather(vw, sapw.last["k_sorted_indices"])
ms = t(lambda: core.attn_fwd(qpw, kpw, vpw, planw, o_rows=sapw.last["q_sorted_indices"]))
flw = 4.0 * D * (sapw.last["q_sizes"].double()[:, :, None] * sapw.last["k_sizes"].double()[:, None, :]
                 * sapw.last["dynamic_map"]).sum().item()
emit(stage="wan.attention_on_kmeans_map", ms=ms, tflops=flw / ms / 1e9, density=float(dw.mean()))
svg1w = wan.WanSVG1Core(Fw, Pw, Hw, D, 0.30, dev)
rows = torch.randint(0, 10000, (64,))
emit(stage="wan.svg1_sparse_core_total", ms=t(lambda: svg1w.sparse_core(qw, kw, vw, rows)))
W_w = svg1w.block_mask.m2
qi = torch.arange(Sw, dtype=torch.int64)
lo = torch.clamp(qi - W_w, min=Pw)
hi = torch.clamp(qi + W_w, max=Sw - 1)
pairs_w = (torch.clamp(hi - lo + 1, min=0) + Pw).sum().item()   # band outside the first frame + first-frame sink
ms = t(lambda: core.attn_fwd(qw, kw, vw, svg1w.block_mask.plan))
emit(stage="wan.svg1_band_attention", ms=ms, W=W_w, tflops=4.0 * D * pairs_w * Hw / ms / 1e9, density=pairs_w / Sw / Sw)
