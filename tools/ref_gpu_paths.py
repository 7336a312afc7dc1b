"""Time the REFERENCE's own GPU block-sparse attention paths on this B200, on the same synthetic inputs as our
kernel, and print ONE JSON object (bench.py runs this as a subprocess under a timeout and embeds the result as
`ref_gpu`; `north_star`: ">= 1.5x the reference's own FlashInfer/Triton block-sparse path on the same B200").

  R-FI  svg.kmeans_utils.dynamic_block_sparse_fwd_flashinfer (svg/kmeans_utils.py:1319-1392): the live SVG2 attention
        call = scratch allocation + VariableBlockSparseAttentionWrapper.plan + .run, every call.  Executed from the
        unmodified reference (baseline/_ref or /root/reference) through tests/golden/ref_import.py; on this image's
        FlashInfer 0.6.x the reference's private-buffer reset no longer exists and is skipped (see ref_import).
        `run_only` re-runs wrapper.run on an existing plan (what a reference that cached its plan would pay).
  R-FX  the reference's compiled flex_attention with its own BlockMask: svg.models.hyvideo.attention
        .prepare_flexattention + Hunyuan_SVGAttn_Processor2_0.sparse_flex_attention (hyvideo/attention.py:30,
        401-403, 527-551; mask_mod hyvideo/utils.py:20-44) — the live SVG1 attention call.

Same inputs go through svgb200 for the ratio and the max-abs difference of the two outputs.  First calls (JIT /
torch.compile) are excluded from the timings and reported separately.  Every section is wrapped so that a failure is
recorded instead of aborting the rest; nothing here is product code.

    python tools/ref_gpu_paths.py [--heads 24] [--skip-flex] [--budget-s 240]
"""
import argparse
import json
import sys
import time
import traceback
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "sparse-videogen_b200"))
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))
import bench  # noqa: E402


def timeit(fn, warm=1, iters=3):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--heads", type=int, default=24)
    ap.add_argument("--skip-flex", action="store_true")
    ap.add_argument("--budget-s", type=float, default=240.0)
    args = ap.parse_args()
    t_start = time.time()
    dev = torch.device("cuda:0")
    H, S, D, F, P, CTX, PLEN = args.heads, bench.S, bench.D, bench.F, bench.P, bench.CTX, bench.PROMPT_LEN
    out = {"heads": H, "S": S, "D": D, "gpu": torch.cuda.get_device_name(0)}
    import ref_import as R

    out["reference_root"] = R.reference_root()
    g = torch.Generator(device=dev).manual_seed(0)
    q, k, v = (torch.randn(1, H, S, D, device=dev, generator=g).to(torch.bfloat16) for _ in range(3))
    from svgb200 import core

    # ------------------------------------------------------------------ SVG2: variable-block map, rho = 0.30
    try:
        import flashinfer

        out["flashinfer_version"] = getattr(flashinfer, "__version__", "?")
        bm, row, col, fl = bench.svg2_map(H, seed=7)
        bmd, rd, cd = bm.to(dev)[None], row.to(dev)[None], col.to(dev)[None]
        t0 = time.time()
        o_ref, how = R.reference_flashinfer_varblock(q, k, v, bmd, rd, cd)
        torch.cuda.synchronize()
        first = time.time() - t0
        ms_call = timeit(lambda: R.reference_flashinfer_varblock(q, k, v, bmd, rd, cd), warm=0, iters=3)
        # run only, on a plan that is kept (the reference re-plans every call)
        fws = torch.empty(128 * 1024 * 1024, device=dev)
        w = flashinfer.sparse.VariableBlockSparseAttentionWrapper(fws, backend="auto")
        w.plan(block_mask_map=bmd[0], block_row_sz=rd[0], block_col_sz=cd[0], num_qo_heads=H, num_kv_heads=H, head_dim=D,
               q_data_type=q.dtype, kv_data_type=k.dtype)
        q3, k3, v3 = (t.reshape(H, S, D) for t in (q, k, v))
        ms_run = timeit(lambda: w.run(q3, k3, v3), warm=1, iters=3)
        del w, fws

        def ours():
            return core.attn_fwd(q, k, v, core.plan_varblock(bmd[0], rd[0], cd[0], S))
        o_ours = ours()
        ms_ours = timeit(ours, warm=1, iters=5)
        out["svg2_varblock"] = {
            "workload": f"QC=400+2 / KC=1000+2 Bernoulli(0.30) map incl. prompt / padding blocks, {H} heads (bench.svg2_map)",
            "reference": "svg.kmeans_utils.dynamic_block_sparse_fwd_flashinfer (svg/kmeans_utils.py:1319-1392): " + how,
            "ref_ms_per_call": ms_call, "ref_ms_run_only": ms_run, "ref_first_call_s": first,
            "ref_tflops": fl / ms_call / 1e9, "ref_tflops_run_only": fl / ms_run / 1e9,
            "ours_ms_per_call_plan_plus_run": ms_ours, "ours_tflops": fl / ms_ours / 1e9,
            "speedup_vs_call": ms_call / ms_ours, "speedup_vs_run_only": ms_run / ms_ours,
            "max_abs_diff": (o_ours.float() - o_ref.float()).abs().max().item(),
        }
        del o_ref, o_ours
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        out["svg2_varblock"] = {"error": repr(e)[:300], "tb": traceback.format_exc()[-800:]}

    # ------------------------------------------------------------------ SVG1: band mask, rho = 0.30
    try:
        if args.skip_flex:
            raise RuntimeError("skipped (--skip-flex)")
        if time.time() - t_start > args.budget_s * 0.5:
            raise RuntimeError("skipped: time budget used by the FlashInfer section")
        A = R.import_model_module("hyvideo", "attention")
        U = R.import_model_module("hyvideo", "utils")
        W, mul = bench.band_width()
        t0 = time.time()
        block_mask = A.prepare_flexattention(1, H, D, torch.bfloat16, dev, CTX, PLEN, F, P, diag_width=mul, multiplier=mul)
        proc = A.Hunyuan_SVGAttn_Processor2_0(layer_idx=0)
        o_ref = proc.sparse_flex_attention(q, k, v, block_mask=block_mask)
        torch.cuda.synchronize()
        first = time.time() - t0
        ms_ref = timeit(lambda: proc.sparse_flex_attention(q, k, v, block_mask=block_mask), warm=1, iters=3)
        fl = 4.0 * D * bench.band_pairs(W) * H
        plan = core.plan_band(core.MASK_HY, F * P, F * P + PLEN, W, H, S, dev)
        o_ours = core.attn_fwd(q, k, v, plan)
        ms_ours = timeit(lambda: core.attn_fwd(q, k, v, plan), warm=1, iters=5)
        out["svg1_band"] = {
            "workload": f"HY band mask W={W} (rho 0.295), {H} heads",
            "reference": "svg.models.hyvideo.attention prepare_flexattention + sparse_flex_attention "
                         "(torch.compile(flex_attention) + BlockMask; hyvideo/attention.py:30,401-403,527-551), unmodified",
            "ref_ms_per_call": ms_ref, "ref_compile_s": first, "ref_tflops": fl / ms_ref / 1e9,
            "ours_ms_per_call": ms_ours, "ours_tflops": fl / ms_ours / 1e9, "speedup": ms_ref / ms_ours,
            "max_abs_diff": (o_ours.float() - o_ref.float()).abs().max().item(),
        }
    except Exception as e:  # noqa: BLE001
        out["svg1_band"] = {"error": repr(e)[:300], "tb": traceback.format_exc()[-800:]}
    out["seconds"] = time.time() - t_start
    print("REF_GPU_JSON " + json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
