"""CUDA-event timings of the pre-attention chain and the Wan block glue at BASELINE shapes, with the reference's
eager (non-_kernels) PyTorch path timed beside it.  JSON lines to gpurun_out/prep.jsonl."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "sparse-videogen_b200"))
from svgb200 import _kernels, core, prep, triton_glue as tg  # noqa: E402

dev = torch.device("cuda:0")
OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)


import os
TAG = os.environ.get("PREP_TAG", "")
FUSED_ONLY = bool(os.environ.get("PREP_FUSED_ONLY"))


def emit(**kw):
    if TAG:
        kw["tag"] = TAG
    if FUSED_ONLY and "fused_qkv" not in kw["op"]:
        return
    print(json.dumps(kw), flush=True)
    with open(OUT / "prep.jsonl", "a") as f:
        f.write(json.dumps(kw) + "\n")


def t(fn, warm=2, iters=5):
    import inspect
    if FUSED_ONLY and "qkv_prep" not in inspect.getsource(fn):
        return 1.0
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


def gbs(nbytes, ms):
    return nbytes / ms / 1e6


# ---- HunyuanVideo 720p single-stream block
S, H, D, txt = 119056, 24, 128, 256
one = S * H * D * 2
qi, ki, vi = (torch.randn(1, S, H * D, device=dev).bfloat16() for _ in range(3))
gq, gk = (torch.randn(D, device=dev).bfloat16() for _ in range(2))
cos, sin = (torch.randn(S - txt, D, device=dev) for _ in range(2))
q = qi.unflatten(2, (H, -1)).transpose(1, 2).contiguous()
k = ki.unflatten(2, (H, -1)).transpose(1, 2).contiguous()
ms = t(lambda: _kernels.rms_norm_forward(q.view(-1, D), gq, 1e-6))
emit(op="hy.rms_norm_inplace(one tensor)", ms=ms, gbs=gbs(2 * one, ms))
ms = t(lambda: _kernels.apply_qk_rope_inplace_cossin_txtlast(q, k, cos, sin, txt))
emit(op="hy.rope_txtlast_inplace(q+k)", ms=ms, gbs=gbs(4 * one, ms))
ms = t(lambda: [x.unflatten(2, (H, -1)).transpose(1, 2).contiguous() for x in (qi, ki, vi)])
emit(op="hy.torch_transpose_contiguous(q,k,v)", ms=ms, gbs=gbs(6 * one, ms))
out = tuple(torch.empty(1, H, S, D, device=dev, dtype=torch.bfloat16) for _ in range(3))
ms = t(lambda: core.qkv_prep(qi, ki, vi, H, out=out, norm=core.NORM_RMS_HEAD, gamma_q=gq, gamma_k=gk, eps=1e-6, rope=1,
                             cos=cos, sin=sin, rope_lo=0, rope_n=S - txt))
emit(op="hy.fused_qkv_prep(transpose+rmsnorm+rope)", ms=ms, gbs=gbs(6 * one, ms), algorithmic_bytes=6 * one)


def stepwise():
    a, b, c = (x.unflatten(2, (H, -1)).transpose(1, 2).contiguous() for x in (qi, ki, vi))
    _kernels.rms_norm_forward(a.view(-1, D), gq, 1e-6)
    _kernels.rms_norm_forward(b.view(-1, D), gk, 1e-6)
    _kernels.apply_qk_rope_inplace_cossin_txtlast(a, b, cos, sin, txt)
    return a, b, c


emit(op="hy.stepwise(torch transpose + in-place norm + in-place rope)", ms=t(stepwise))


def eager():  # the reference's ENABLE_FAST_KERNEL=False branch (hyvideo/attention.py:189-222) with diffusers' formulas
    a, b, c = (x.unflatten(2, (H, -1)).transpose(1, 2).contiguous() for x in (qi, ki, vi))
    a = torch.nn.functional.rms_norm(a, [D], gq, 1e-6)
    b = torch.nn.functional.rms_norm(b, [D], gk, 1e-6)

    def rope(x):
        img = x[:, :, :-txt]
        xr, xim = img.reshape(*img.shape[:-1], -1, 2).unbind(-1)
        rot = torch.stack([-xim, xr], dim=-1).flatten(3)
        return torch.cat([(img.float() * cos + rot.float() * sin).to(x.dtype), x[:, :, -txt:]], dim=2)
    return rope(a), rope(b), c


del q, k
torch.cuda.empty_cache()
emit(op="hy.eager_torch_chain(reference fallback path)", ms=t(eager, warm=1, iters=2))
del qi, ki, vi, out
torch.cuda.empty_cache()

# ---- Wan 2.1 14B 720p block glue: rows = 75600, hidden = 5120
Sw, N, Hw = 75600, 5120, 40
x = torch.randn(1, Sw, N, device=dev).bfloat16()
attn = torch.randn(1, Sw, N, device=dev).bfloat16()
scale, shift, gate = (torch.randn(1, 1, N, device=dev) for _ in range(3))
w = torch.randn(N, device=dev).bfloat16()
xb = Sw * N * 2
ms = t(lambda: tg.layernorm_modulate_forward(x, None, None, 1e-6, scale, shift))
emit(op="wan.fused_layernorm_modulate(bf16->bf16)", ms=ms, gbs=gbs(2 * xb, ms))
ms = t(lambda: tg.triton_modulate_shift_forward(tg.triton_layernorm_forward(x, None, None, 1e-6, False), scale, shift,
                                                torch.bfloat16))
emit(op="wan.layernorm(fp32 out) + modulate (reference call sequence)", ms=ms, gbs=gbs(xb + 2 * xb * 2 + xb, ms))
ms = t(lambda: tg.triton_modulate_gate_residual_forward(x, attn, gate, torch.bfloat16))
emit(op="wan.gate_residual", ms=ms, gbs=gbs(3 * xb, ms))
ms = t(lambda: tg.triton_rmsnorm_forward(x, w, 1e-6))
emit(op="wan.rmsnorm_hidden", ms=ms, gbs=gbs(2 * xb, ms))
ln = torch.nn.LayerNorm(N, eps=1e-6, elementwise_affine=False)
ms = t(lambda: (ln(x.float()) * (1 + scale) + shift).type_as(x), warm=1, iters=3)
emit(op="wan.eager_torch layernorm+modulate (reference fallback)", ms=ms)
fr, fi = torch.randn(Sw, 64, device=dev), torch.randn(Sw, 64, device=dev)
qi, ki, vi = x, attn, torch.randn(1, Sw, N, device=dev).bfloat16()
out = tuple(torch.empty(1, Hw, Sw, 128, device=dev, dtype=torch.bfloat16) for _ in range(3))
ms = t(lambda: core.qkv_prep(qi, ki, vi, Hw, out=out, norm=core.NORM_RMS_HIDDEN, gamma_q=w, gamma_k=w, eps=1e-6, rope=2,
                             cos=fr, sin=fi, rope_lo=0, rope_n=Sw))
emit(op="wan.fused_qkv_prep(rmsnorm_hidden+transpose+complex rope)", ms=ms, gbs=gbs(6 * xb, ms))


def wan_stepwise():
    a = tg.triton_rmsnorm_forward(qi, w, 1e-6)
    b = tg.triton_rmsnorm_forward(ki, w, 1e-6)
    a, b, c = (z.unflatten(2, (Hw, -1)).transpose(1, 2).contiguous() for z in (a, b, vi))
    _kernels.apply_qk_rope_inplace_cossin_complex(a, b, fr, fi, 0)
    return a, b, c


emit(op="wan.stepwise(rmsnorm + torch transpose + in-place complex rope)", ms=t(wan_stepwise))
