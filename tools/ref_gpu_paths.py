"""Time the REFERENCE's own GPU block-sparse attention paths on this B200, on the same synthetic inputs
as our kernel (R-FI, R-FX).  Measurement tool only — nothing here is product code.

  R-FI  svg.kmeans_utils.dynamic_block_sparse_fwd_flashinfer = FlashInfer VariableBlockSparseAttentionWrapper
        (svg/kmeans_utils.py:1319-1392), plan + run as the reference calls it every step, and run alone.
        The image ships flashinfer 0.6.x (the reference vendors 0.2.10 + a patch; same wrapper API).
  R-FX  torch.compile(flex_attention) with the reference HY BlockMask
        (svg/models/hyvideo/attention.py:30,401-403,527-551; mask_mod hyvideo/utils.py:20-44).

Writes JSON lines to gpurun_out/ref_gpu.jsonl.  Each section is wrapped so a failure (JIT unavailable
offline, OOM...) is recorded instead of aborting the rest.
"""
import json
import math
import os
import sys
import time
import traceback
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "sparse-videogen_b200"))
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)
dev = torch.device("cuda:0")
H = int(os.environ.get("REF_H", 24))
S, D, F, P, CTX, PLEN = bench.S, bench.D, bench.F, bench.P, bench.CTX, bench.PROMPT_LEN


def emit(**kw):
    print(json.dumps(kw), flush=True)
    with open(OUT / "ref_gpu.jsonl", "a") as f:
        f.write(json.dumps(kw) + "\n")


def timeit(fn, warm=2, iters=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


q, k, v = (torch.randn(1, H, S, D, device=dev, dtype=torch.bfloat16) for _ in range(3))


def varblock_inputs(QC, KC, rho, seed=0):
    g = torch.Generator().manual_seed(seed)

    def sizes(n):
        b = torch.full((H, n), S // n, dtype=torch.int32)
        b[:, : S - (S // n) * n] += 1
        return b
    row, col = sizes(QC), sizes(KC)
    bm = torch.rand(H, QC, KC, generator=g) < rho
    bm[:, :, 0] = True
    fl = 4.0 * D * (row.double()[:, :, None] * col.double()[:, None, :] * bm).sum().item()
    return bm, row, col, fl


# ------------------------------------------------------------------------------------------ ours
try:
    from svgb200 import core

    for QC, KC in ((400, 1000), (465, 931)):
        bm, row, col, fl = varblock_inputs(QC, KC, 0.3)
        bmd, rd, cd = bm.to(dev), row.to(dev), col.to(dev)

        def ours():
            plan = core.plan_varblock(bmd, rd, cd, S)
            return core.attn_fwd(q, k, v, plan)
        ms = timeit(ours)
        emit(path="ours.varblock(plan+run)", QC=QC, KC=KC, ms=ms, tflops=fl / ms / 1e9)
except Exception as e:  # noqa: BLE001
    emit(path="ours.varblock", error=repr(e))

# ------------------------------------------------------------------------------------------ R-FI
try:
    import flashinfer

    emit(path="flashinfer", version=getattr(flashinfer, "__version__", "?"))
    for QC, KC in ((400, 1000), (465, 931)):
        bm, row, col, fl = varblock_inputs(QC, KC, 0.3)
        bmd, rd, cd = bm.to(dev), row.to(dev), col.to(dev)
        float_ws = torch.empty(128 * 1024 * 1024, device=dev)            # kmeans_utils.py:1358
        vec_idx = torch.empty(1024 * 1024 * 1024, device=dev)             # :1359 (4 GiB fp32 scratch)
        wrapper = flashinfer.sparse.VariableBlockSparseAttentionWrapper(float_ws, backend="auto")
        try:  # the reference enlarges the index scratch (kmeans_utils.py:1361-1366); attribute names moved in 0.6.x
            wrapper.reset_workspace_buffer(float_workspace_buffer=wrapper._float_workspace_buffer,
                                           int_workspace_buffer=wrapper._int_workspace_buffer,
                                           vector_sparse_indices_buffer=vec_idx,
                                           vector_sparse_indptr_buffer=wrapper._vector_sparse_indptr_buffer)
        except Exception as e:  # noqa: BLE001
            emit(path="R-FI.note", note="reset_workspace_buffer unavailable in this flashinfer: " + repr(e)[:200])
        q3, k3, v3 = (t.reshape(H, S, D) for t in (q, k, v))

        def plan():
            wrapper.plan(block_mask_map=bmd, block_row_sz=rd, block_col_sz=cd, num_qo_heads=H, num_kv_heads=H,
                         head_dim=D, q_data_type=q.dtype, kv_data_type=k.dtype)

        t0 = time.time()
        plan()
        o = wrapper.run(q3, k3, v3)
        torch.cuda.synchronize()
        emit(path="R-FI.first_call_s", QC=QC, KC=KC, seconds=time.time() - t0)
        ms_run = timeit(lambda: wrapper.run(q3, k3, v3), warm=1, iters=3)

        def both():
            plan()
            return wrapper.run(q3, k3, v3)
        ms_both = timeit(both, warm=1, iters=3)
        emit(path="R-FI.varblock", QC=QC, KC=KC, ms_run=ms_run, ms_plan_run=ms_both, tflops_run=fl / ms_run / 1e9,
             tflops_plan_run=fl / ms_both / 1e9)
        try:
            from svgb200 import core
            ours_o = core.attn_fwd(q, k, v, core.plan_varblock(bmd, rd, cd, S))
            diff = (ours_o.view(H, S, D).float() - o.float()).abs().max().item()
            emit(path="R-FI.vs_ours_max_abs_diff", QC=QC, KC=KC, diff=diff)
        except Exception as e:  # noqa: BLE001
            emit(path="R-FI.compare", error=repr(e))
        del wrapper, float_ws, vec_idx
        torch.cuda.empty_cache()
except Exception as e:  # noqa: BLE001
    emit(path="R-FI", error=repr(e), tb=traceback.format_exc()[-1500:])

# ------------------------------------------------------------------------------------------ R-FX
try:
    if os.environ.get("REF_SKIP_FX"):
        raise RuntimeError("skipped (REF_SKIP_FX)")
    from torch.nn.attention.flex_attention import create_block_mask, flex_attention

    W, mul = bench.band_width()
    real_length = F * P + PLEN

    def temporal_mask_mod(b, h, q_idx, kv_idx):   # hyvideo/utils.py:29-42
        real_mask = (kv_idx < real_length) & (q_idx < real_length)
        fake_mask = (kv_idx >= real_length) & (q_idx >= real_length)
        temporal_head_mask = torch.abs(q_idx - kv_idx) < W
        text_column_mask = (F * P <= kv_idx) & (kv_idx < real_length)
        text_row_mask = (F * P <= q_idx) & (q_idx < real_length)
        return (real_mask & (temporal_head_mask | text_column_mask | text_row_mask)) | fake_mask

    t0 = time.time()
    block_mask = create_block_mask(temporal_mask_mod, None, None, S, S, device=dev, _compile=True)
    flex = torch.compile(flex_attention, dynamic=False)                  # hyvideo/attention.py:30
    o = flex(q, k, v, block_mask=block_mask)
    torch.cuda.synchronize()
    emit(path="R-FX.compile_s", seconds=time.time() - t0)
    ms = timeit(lambda: flex(q, k, v, block_mask=block_mask), warm=1, iters=3)
    fl = 4.0 * D * bench.band_pairs(W) * H
    emit(path="R-FX.flex_attention_band", W=W, ms=ms, tflops=fl / ms / 1e9)
    try:
        from svgb200 import core
        plan = core.plan_band(core.MASK_HY, F * P, real_length, W, H, S, dev)
        ours_o = core.attn_fwd(q, k, v, plan)
        emit(path="R-FX.vs_ours_max_abs_diff", diff=(ours_o.float() - o.float()).abs().max().item())
        ms_o = timeit(lambda: core.attn_fwd(q, k, v, plan))
        emit(path="ours.band", ms=ms_o, tflops=fl / ms_o / 1e9, speedup_vs_flex=ms / ms_o)
    except Exception as e:  # noqa: BLE001
        emit(path="R-FX.compare", error=repr(e))
except Exception as e:  # noqa: BLE001
    emit(path="R-FX", error=repr(e), tb=traceback.format_exc()[-1500:])
