"""Low-level torch-facing wrappers over the C ABI: pointer/stride plumbing, workspace caching,
argument checks.  Everything here runs the CUDA library; there is no eager fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from ._lib import Plan, SvgbError, check, lib

MASK_NONE, MASK_HY, MASK_WAN, MASK_COG = 0, 1, 2, 3

# launches issued through this module since import (bench.py reports it as gpu_launches)
launch_count = 0


def _bump(n=1):
    global launch_count
    launch_count += n


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return 0
    if t.dtype == torch.float16:
        return 1
    raise SvgbError(f"unsupported dtype {t.dtype}: svgb200 kernels take bfloat16 or float16")


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise SvgbError("svgb200 operators need CUDA tensors (no CPU fallback)")


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


_ws_cache: dict = {}


def workspace(key, nbytes: int, device) -> torch.Tensor:
    """Cached uint8 scratch buffer (stream-ordered reuse on the current stream)."""
    k = (key, str(device))
    buf = _ws_cache.get(k)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _ws_cache[k] = buf
    return buf


def device_check():
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    check(lib().svgb_device_check(C.byref(a), C.byref(b), C.byref(c)), "svgb_device_check")
    return a.value, b.value, c.value


# ----------------------------------------------------------------------------------------------
# plans
# ----------------------------------------------------------------------------------------------
@dataclass
class AttnPlan:
    """Device-resident work list + its host descriptor."""
    desc: Plan
    ws: torch.Tensor

    @property
    def S(self):
        return self.desc.S


def plan_varblock(block_map: torch.Tensor, row_sz: torch.Tensor, col_sz: torch.Tensor, S: int,
                  ws: Optional[torch.Tensor] = None, gather: bool = False, shared_scratch: bool = False) -> AttnPlan:
    """block_map [BH,QC,KC] bool/uint8, row_sz [BH,QC], col_sz [BH,KC] (any int dtype) on GPU.
    gather=True lowers the map for the row-gather kernel path (exactly-full chunks, optional fused
    permutation through attn_fwd(..., q_rows=, kv_rows=)).

    Workspace ownership: by default every plan gets its OWN device work list (torch's caching allocator makes the
    per-call allocation free of cudaMalloc / syncs), so two plans of one shape can be held and executed in any
    order or on different streams.  shared_scratch=True reuses one shape-keyed scratch buffer instead: only for
    callers that execute the plan immediately, on the same stream, before building the next one of that shape."""
    _need_cuda(block_map, row_sz, col_sz)
    BH, QC, KC = block_map.shape
    m = block_map.contiguous()
    m = m.view(torch.uint8) if m.dtype == torch.bool else m.to(torch.uint8)
    r = row_sz.reshape(BH, QC).to(torch.int32).contiguous()
    c = col_sz.reshape(BH, KC).to(torch.int32).contiguous()
    nbytes = C.c_size_t()
    check(lib().svgb_attn_plan_varblock_bytes(BH, S, QC, KC, C.byref(nbytes)), "plan_varblock_bytes")
    if ws is None:
        ws = (workspace(("vb", BH, S, QC, KC), nbytes.value, m.device) if shared_scratch
              else torch.empty(nbytes.value, dtype=torch.uint8, device=m.device))
    desc = Plan()
    fn = lib().svgb_attn_plan_varblock_gather if gather else lib().svgb_attn_plan_varblock
    check(fn(m.data_ptr(), r.data_ptr(), c.data_ptr(), BH, S, QC, KC, ws.data_ptr(), ws.numel(), C.byref(desc),
             _stream(m)), "svgb_attn_plan_varblock")
    _bump()
    return AttnPlan(desc, ws)


def plan_band(mask_mode: int, m0: int, m1: int, m2: int, BH: int, S: int, device) -> AttnPlan:
    nbytes = C.c_size_t()
    check(lib().svgb_attn_plan_band_bytes(S, C.byref(nbytes)), "plan_band_bytes")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=device)  # owned by the plan (long-lived)
    desc = Plan()
    check(lib().svgb_attn_plan_band(mask_mode, m0, m1, m2, BH, S, ws.data_ptr(), ws.numel(),
                                    C.byref(desc), torch.cuda.current_stream(device).cuda_stream),
          "svgb_attn_plan_band")
    _bump()
    return AttnPlan(desc, ws)


# ----------------------------------------------------------------------------------------------
# attention
# ----------------------------------------------------------------------------------------------
def attn_fwd(q, k, v, plan: AttnPlan, *, layout: str = "bhsd", o_rows: Optional[torch.Tensor] = None,
             q_rows: Optional[torch.Tensor] = None, kv_rows: Optional[torch.Tensor] = None,
             return_lse: bool = False, sm_scale: Optional[float] = None, out: Optional[torch.Tensor] = None):
    """layout 'bhsd': q,k,v [B,H,S,D] contiguous;  'shd': [S,H,D] contiguous (ops API)."""
    _need_cuda(q, k, v)
    if not (q.is_contiguous() and k.is_contiguous() and v.is_contiguous()):
        raise SvgbError("q, k, v must be contiguous")
    if not (q.dtype == k.dtype == v.dtype) or not (q.shape == k.shape == v.shape):
        raise SvgbError("q, k, v must share dtype and shape")
    if layout == "bhsd":
        B, H, S, D = q.shape
        BH, rs, hs = B * H, D, S * D
    elif layout == "shd":
        S, H, D = q.shape
        BH, rs, hs = H, H * D, D
    else:
        raise SvgbError(f"unknown layout {layout}")
    o = torch.empty_like(q) if out is None else out
    lse = torch.empty(BH, S, dtype=torch.float32, device=q.device) if return_lse else None
    if o_rows is not None:
        o_rows = o_rows.reshape(BH, S).to(torch.int32).contiguous()
    scale = float(D) ** -0.5 if sm_scale is None else float(sm_scale)
    if plan.desc.kind == 3:
        if q_rows is not None:
            q_rows = q_rows.reshape(BH, S).to(torch.int32).contiguous()
        if kv_rows is not None:
            kv_rows = kv_rows.reshape(BH, S).to(torch.int32).contiguous()
        check(lib().svgb_attn_fwd_gather(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), _p(lse),
                                         _p(q_rows), _p(kv_rows), _p(o_rows), _dt(q), BH, S, D, rs, hs, rs, hs,
                                         scale, C.byref(plan.desc), plan.ws.data_ptr(), _stream(q)),
              "svgb_attn_fwd_gather")
    else:
        if q_rows is not None or kv_rows is not None:
            raise SvgbError("q_rows / kv_rows need a gather plan (plan_varblock(..., gather=True))")
        check(lib().svgb_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), _p(lse), _p(o_rows),
                                  _dt(q), BH, S, D, rs, hs, rs, hs, scale, C.byref(plan.desc),
                                  plan.ws.data_ptr(), _stream(q)), "svgb_attn_fwd")
    _bump()
    return (o, lse) if return_lse else o


def quantize_e4m3(x: torch.Tensor):
    """Per-head absmax quantisation of a 16-bit [..., S, D] tensor to fp8 e4m3 bytes.
    Returns (x8 uint8 same shape, scale fp32 [BH]) with x ~= x8.view(float8_e4m3fn) * scale[h]."""
    _need_cuda(x)
    S, D = x.shape[-2], x.shape[-1]
    xc = x.contiguous()
    BH = xc.numel() // (S * D)
    x8 = torch.empty(xc.shape, dtype=torch.uint8, device=x.device)
    scale = torch.empty(BH, dtype=torch.float32, device=x.device)
    check(lib().svgb_quantize_e4m3(xc.data_ptr(), _dt(x), x8.data_ptr(), scale.data_ptr(), BH, S, D, _stream(x)),
          "svgb_quantize_e4m3")
    _bump(3)
    return x8, scale


def attn_fwd_fp8(q8, k8, v8, q_scale, k_scale, v_scale, plan: AttnPlan, *, o_rows: Optional[torch.Tensor] = None,
                 return_lse: bool = False, sm_scale: Optional[float] = None):
    """FP8 (e4m3) block-sparse attention: q8,k8,v8 uint8 [B,H,S,128] + per-head scales -> bf16 [B,H,S,128]."""
    _need_cuda(q8, k8, v8, q_scale, k_scale, v_scale)
    if not (q8.dtype == k8.dtype == v8.dtype == torch.uint8) or not (q8.shape == k8.shape == v8.shape):
        raise SvgbError("q8, k8, v8 must be uint8 tensors of one shape")
    B, H, S, D = q8.shape
    BH = B * H
    o = torch.empty(B, H, S, D, dtype=torch.bfloat16, device=q8.device)
    lse = torch.empty(BH, S, dtype=torch.float32, device=q8.device) if return_lse else None
    if o_rows is not None:
        o_rows = o_rows.reshape(BH, S).to(torch.int32).contiguous()
    scale = float(D) ** -0.5 if sm_scale is None else float(sm_scale)
    # temporaries are bound to locals that outlive the launch call (a dropped .contiguous() copy could be recycled
    # by the allocator for the next one and alias two operands)
    q8c, k8c, v8c = q8.contiguous(), k8.contiguous(), v8.contiguous()
    sqc, skc, svc = (t.to(torch.float32).contiguous() for t in (q_scale, k_scale, v_scale))
    check(lib().svgb_attn_fwd_fp8(q8c.data_ptr(), k8c.data_ptr(), v8c.data_ptr(), sqc.data_ptr(), skc.data_ptr(),
                                  svc.data_ptr(), o.data_ptr(), _p(lse), _p(o_rows), BH, S, D,
                                  D, S * D, D, S * D, scale, C.byref(plan.desc), plan.ws.data_ptr(), _stream(q8)),
          "svgb_attn_fwd_fp8")
    _bump()
    return (o, lse) if return_lse else o


def density(block_map, row_sz, col_sz):
    _need_cuda(block_map, row_sz, col_sz)
    BH, QC, KC = block_map.shape
    m = block_map.contiguous()
    m = m.view(torch.uint8) if m.dtype == torch.bool else m.to(torch.uint8)
    r = row_sz.reshape(BH, QC).to(torch.int32).contiguous()
    c = col_sz.reshape(BH, KC).to(torch.int32).contiguous()
    out = torch.empty(BH, dtype=torch.float32, device=m.device)
    check(lib().svgb_density(m.data_ptr(), r.data_ptr(), c.data_ptr(), BH, QC, KC, out.data_ptr(), _stream(m)),
          "svgb_density")
    _bump()
    return out


# ----------------------------------------------------------------------------------------------
# layout transforms
# ----------------------------------------------------------------------------------------------
def argsort_labels(labels: torch.Tensor, K: int):
    """labels int [BH,S] in [0,K) -> (perm int32 [BH,S] stable ascending, counts int32 [BH,K])."""
    _need_cuda(labels)
    BH, S = labels.shape
    lab = labels.to(torch.int32).contiguous()
    nbytes = C.c_size_t()
    check(lib().svgb_argsort_labels_bytes(BH, S, K, C.byref(nbytes)), "argsort_labels_bytes")
    ws = workspace(("sort", BH, S, K), nbytes.value, lab.device)
    perm = torch.empty(BH, S, dtype=torch.int32, device=lab.device)
    counts = torch.empty(BH, K, dtype=torch.int32, device=lab.device)
    check(lib().svgb_argsort_labels(lab.data_ptr(), BH, S, K, perm.data_ptr(), counts.data_ptr(),
                                    ws.data_ptr(), ws.numel(), _stream(lab)), "svgb_argsort_labels")
    _bump(3)
    return perm, counts


def _rows_op(fn, name, x, perm):
    _need_cuda(x, perm)
    if x.element_size() != 2:
        raise SvgbError("permute kernels move 16-bit elements")
    shape = x.shape
    D = shape[-1]
    S = shape[-2]
    xf = x.contiguous().view(-1, S, D)
    BH = xf.shape[0]
    p = perm.reshape(BH, S).to(torch.int32).contiguous()
    out = torch.empty_like(xf)
    check(fn(xf.data_ptr(), p.data_ptr(), out.data_ptr(), BH, S, D, _stream(x)), name)
    _bump()
    return out.view(shape)


def permute_gather(x, perm):
    return _rows_op(lib().svgb_permute_gather, "svgb_permute_gather", x, perm)


def permute_scatter(x, perm):
    return _rows_op(lib().svgb_permute_scatter, "svgb_permute_scatter", x, perm)


def head_placement(ins, outs, best_mask_idx, ctx, F, P, *, text_first=False, inverse=False):
    """ins / outs: lists (1..3) of [cfg,H,S,D] contiguous 16-bit tensors; writes outs in place."""
    _need_cuda(*ins, *outs, best_mask_idx)
    n = len(ins)
    cfg, H, S, D = ins[0].shape
    for t in list(ins) + list(outs):
        if not t.is_contiguous() or t.shape != ins[0].shape or t.element_size() != 2:
            raise SvgbError("placement tensors must be contiguous 16-bit [cfg,H,S,D] of one shape")
    idx = best_mask_idx.reshape(cfg * H).to(torch.int32).contiguous()
    a_in = (C.c_void_p * n)(*[t.data_ptr() for t in ins])
    a_out = (C.c_void_p * n)(*[t.data_ptr() for t in outs])
    check(lib().svgb_head_placement(a_in, a_out, n, idx.data_ptr(), cfg * H, S, D, ctx, F, P,
                                    1 if text_first else 0, 1 if inverse else 0, _stream(ins[0])),
          "svgb_head_placement")
    _bump()
    return outs


# ----------------------------------------------------------------------------------------------
# flash k-means
# ----------------------------------------------------------------------------------------------
def _km_ws(BH, N, K, D, device):
    nbytes = C.c_size_t()
    check(lib().svgb_kmeans_bytes(BH, N, K, D, C.byref(nbytes)), "svgb_kmeans_bytes")
    return workspace(("km", BH, N, K, D), nbytes.value, device)


def row_sqnorm(x, round_result=True):
    """x [BH,N,D] 16-bit -> fp32 [BH,N]"""
    _need_cuda(x)
    BH, N, D = x.shape
    xc = x.contiguous()
    out = torch.empty(BH, N, dtype=torch.float32, device=x.device)
    check(lib().svgb_row_sqnorm(xc.data_ptr(), out.data_ptr(), BH, N, D, _dt(x), 1 if round_result else 0,
                                _stream(x)), "svgb_row_sqnorm")
    _bump()
    return out


def kmeans_assign(x, c, x_sq):
    _need_cuda(x, c, x_sq)
    BH, N, D = x.shape
    K = c.shape[1]
    xc, cc, sqc = x.contiguous(), c.contiguous(), x_sq.to(torch.float32).contiguous()
    ws = _km_ws(BH, N, K, D, x.device)
    labels = torch.empty(BH, N, dtype=torch.int32, device=x.device)
    check(lib().svgb_kmeans_assign(xc.data_ptr(), cc.data_ptr(), sqc.data_ptr(), labels.data_ptr(),
                                   BH, N, K, D, _dt(x), ws.data_ptr(), ws.numel(), _stream(x)),
          "svgb_kmeans_assign")
    _bump(2)
    return labels


def kmeans_update(x, labels, c_old):
    _need_cuda(x, labels, c_old)
    BH, N, D = x.shape
    K = c_old.shape[1]
    ws = _km_ws(BH, N, K, D, x.device)
    xc, lc, cc = x.contiguous(), labels.to(torch.int32).contiguous(), c_old.contiguous()
    c_new = torch.empty_like(cc)
    counts = torch.empty(BH, K, dtype=torch.int32, device=x.device)
    shift = torch.zeros(1, dtype=torch.float32, device=x.device)
    check(lib().svgb_kmeans_update(xc.data_ptr(), lc.data_ptr(),
                                   cc.data_ptr(), c_new.data_ptr(), counts.data_ptr(),
                                   shift.data_ptr(), BH, N, K, D, _dt(x), ws.data_ptr(), ws.numel(), _stream(x)),
          "svgb_kmeans_update")
    _bump(5)
    return c_new, counts, shift


def kmeans_run(x, init_centroids, max_iters, tol=1e-4, want_perm=False):
    """Whole Lloyd loop on the device (no host sync).  Returns labels int32 [BH,N], centroids [BH,K,D],
    counts int32 [BH,K], n_iter int32[1] (device) and, with want_perm, the stable argsort of the labels int32 [BH,N]
    (the member lists of the last centroid update).  x may be a view whose heads are further apart than N*D (the video
    part of a [H, S, D] tensor): it is clustered in place, without a packing copy."""
    _need_cuda(x, init_centroids)
    BH, N, D = x.shape
    K = init_centroids.shape[1]
    ws = _km_ws(BH, N, K, D, x.device)
    labels = torch.empty(BH, N, dtype=torch.int32, device=x.device)
    cents = torch.empty(BH, K, D, dtype=x.dtype, device=x.device)
    counts = torch.empty(BH, K, dtype=torch.int32, device=x.device)
    n_iter = torch.zeros(1, dtype=torch.int32, device=x.device)
    perm = torch.empty(BH, N, dtype=torch.int32, device=x.device) if want_perm else None
    strided_ok = x.stride(2) == 1 and x.stride(1) == D and (BH == 1 or (x.stride(0) >= N * D and x.stride(0) % 8 == 0))
    xc = x if strided_ok else x.contiguous()
    head_stride = xc.stride(0) if BH > 1 else N * D
    ic = init_centroids.contiguous()
    check(lib().svgb_kmeans_run_sorted(xc.data_ptr(), head_stride, ic.data_ptr(), BH, N, K, D,
                                       _dt(x), int(max_iters), float(tol), labels.data_ptr(), cents.data_ptr(),
                                       counts.data_ptr(), n_iter.data_ptr(), perm.data_ptr() if want_perm else None,
                                       ws.data_ptr(), ws.numel(), _stream(x)),
          "svgb_kmeans_run_sorted")
    _bump(2 + 7 * int(max_iters))
    if want_perm:
        return labels, cents, counts, n_iter, perm
    return labels, cents, counts, n_iter


def dynamic_map(qc, kc, k_sizes, top_p, preserve):
    """qc [BH,QC,D], kc [BH,KC,D] 16-bit, k_sizes int [BH,KC] -> bool [BH,QC,KC]"""
    _need_cuda(qc, kc, k_sizes)
    BH, QC, D = qc.shape
    KC = kc.shape[1]
    out = torch.empty(BH, QC, KC, dtype=torch.uint8, device=qc.device)
    qcc, kcc, ksc = qc.contiguous(), kc.contiguous(), k_sizes.to(torch.int32).contiguous()
    check(lib().svgb_dynamic_map(qcc.data_ptr(), kcc.data_ptr(),
                                 ksc.data_ptr(), BH, QC, KC, D, _dt(qc),
                                 float(top_p), int(preserve), out.data_ptr(), _stream(qc)), "svgb_dynamic_map")
    _bump()
    return out.view(torch.bool)


def sample_mse(q, k, v, rows, layout, ctx, F, P):
    """q,k,v [BH,S,D]; rows int [n] (<=128) -> fp32 [2, BH] (0 = spatial mask, 1 = temporal mask)"""
    _need_cuda(q, k, v, rows)
    BH, S, D = q.shape
    n = rows.numel()
    nbytes = C.c_size_t()
    check(lib().svgb_sample_mse_bytes(BH, S, D, n, C.byref(nbytes)), "svgb_sample_mse_bytes")
    ws = workspace(("smse", BH, S, D), nbytes.value, q.device)
    out = torch.empty(2, BH, dtype=torch.float32, device=q.device)
    r = rows.to(torch.int32).contiguous()
    qc_, kc_, vc_ = q.contiguous(), k.contiguous(), v.contiguous()
    check(lib().svgb_sample_mse(qc_.data_ptr(), kc_.data_ptr(), vc_.data_ptr(),
                                r.data_ptr(), n, BH, S, D, _dt(q), int(layout), ctx, F, P, out.data_ptr(),
                                ws.data_ptr(), ws.numel(), _stream(q)), "svgb_sample_mse")
    _bump(7)
    return out


def selftest_tile(q, k, v, p_scale=0.0625):
    """q,k,v: [128, D] -> (S fp32 [128,128], O fp32 [128,D])"""
    _need_cuda(q, k, v)
    D = q.shape[-1]
    s = torch.empty(128, 128, dtype=torch.float32, device=q.device)
    o = torch.empty(128, D, dtype=torch.float32, device=q.device)
    check(lib().svgb_selftest_tile(q.data_ptr(), k.data_ptr(), v.data_ptr(), s.data_ptr(), o.data_ptr(),
                                   D, _dt(q), p_scale, _stream(q)), "svgb_selftest_tile")
    _bump()
    return s, o


# ---- pre-attention chain (include/svgb200.h: svgb_rms_norm / svgb_layer_norm / svgb_qk_rope / svgb_qkv_prep)
ROPE_TXT_FIRST, ROPE_TXT_LAST, ROPE_COMPLEX_TXT_FIRST = 0, 1, 2
NORM_NONE, NORM_RMS_HEAD, NORM_LAYER, NORM_RMS_HIDDEN = 0, 1, 2, 3


def _dt3(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return 3
    return _dt(t)


def _contig(*ts):
    for t in ts:
        if t is not None and not t.is_contiguous():
            raise SvgbError("tensor must be contiguous")


def rms_norm_(x: torch.Tensor, gamma: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """In place on x [m, n] (any leading shape; last dim n in {32,64,128,256})."""
    _need_cuda(x, gamma)
    _contig(x, gamma)
    if gamma.dim() != 1 or gamma.shape[0] != x.shape[-1] or gamma.dtype != x.dtype:
        raise SvgbError("gamma must be [n] with the dtype of x")
    n = x.shape[-1]
    check(lib().svgb_rms_norm(x.data_ptr(), gamma.data_ptr(), x.numel() // n, n, float(eps), _dt(x), _stream(x)),
          "svgb_rms_norm")
    _bump()
    return x


def layer_norm_(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor) -> torch.Tensor:
    _need_cuda(x, gamma, beta)
    _contig(x, gamma, beta)
    n = x.shape[-1]
    if gamma.shape != (n,) or beta.shape != (n,) or gamma.dtype != x.dtype or beta.dtype != x.dtype:
        raise SvgbError("gamma / beta must be [n] with the dtype of x")
    check(lib().svgb_layer_norm(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), x.numel() // n, n, _dt(x),
                                _stream(x)), "svgb_layer_norm")
    _bump()
    return x


def qk_rope_(q: torch.Tensor, k: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, len_text: int, mode: int):
    """In place on q [B,Hq,S,D], k [B,Hk,S,D]; cos/sin float32 [S-len_text, D] (or [.., D/2] for the complex mode)."""
    _need_cuda(q, k, cos, sin)
    _contig(q, k, cos, sin)
    if q.dim() != 4 or k.dim() != 4 or cos.dim() != 2 or sin.dim() != 2:
        raise SvgbError("q, k must be 4-D and cos, sin 2-D")
    if cos.dtype != torch.float32 or sin.dtype != torch.float32:
        raise SvgbError("cos / sin caches must be float32")
    B, Hq, S, D = q.shape
    if k.shape[0] != B or k.shape[2] != S or k.shape[3] != D or k.dtype != q.dtype:
        raise SvgbError("q and k must agree in batch, sequence, head_dim and dtype")
    valid = S - int(len_text)
    width = D // 2 if mode == ROPE_COMPLEX_TXT_FIRST else D
    if tuple(cos.shape) != (valid, width) or tuple(sin.shape) != (valid, width):
        raise SvgbError(f"cos / sin must be [{valid}, {width}], got {tuple(cos.shape)} / {tuple(sin.shape)}")
    check(lib().svgb_qk_rope(q.data_ptr(), k.data_ptr(), cos.data_ptr(), sin.data_ptr(), B, Hq, k.shape[1], S, D,
                             int(len_text), int(mode), _dt(q), _stream(q)), "svgb_qk_rope")
    _bump()
    return q, k


def qkv_prep(q_in, k_in, v_in, heads: int, *, out=None, out_row0: int = 0, out_rows: Optional[int] = None,
             norm: int = NORM_NONE, gamma_q=None, gamma_k=None, beta_q=None, beta_k=None, eps: float = 1e-6,
             rope: int = 0, cos=None, sin=None, rope_lo: int = 0, rope_n: Optional[int] = None):
    """[B, S_in, H*D] x3 -> (q, k, v) [B, H, S_out, D] rows [out_row0, out_row0+S_in) in one pass (transpose + QK norm
    + RoPE).  q_in/k_in/v_in may be views of one packed projection (last dim contiguous)."""
    _need_cuda(q_in, k_in, v_in, gamma_q, gamma_k, beta_q, beta_k, cos, sin)
    B, S_in, HD = q_in.shape
    D = HD // heads
    for t in (q_in, k_in, v_in):
        if t.shape != q_in.shape or t.dtype != q_in.dtype or t.stride(2) != 1 or t.stride() != q_in.stride():
            raise SvgbError("q_in, k_in, v_in must share shape, dtype and strides (last dim contiguous)")
    S_out = out_rows if out_rows is not None else (out[0].shape[2] if out is not None else S_in)
    if out is None:
        out = tuple(torch.empty(B, heads, S_out, D, dtype=q_in.dtype, device=q_in.device) for _ in range(3))
    for t in out:
        if tuple(t.shape) != (B, heads, S_out, D) or not t.is_contiguous() or t.dtype != q_in.dtype:
            raise SvgbError("outputs must be contiguous [B, H, S_out, D] of the input dtype")
    if out_row0 + S_in > S_out:
        raise SvgbError("out_row0 + S_in exceeds the output rows")
    if rope:
        if cos is None or sin is None or cos.dtype != torch.float32 or sin.dtype != torch.float32:
            raise SvgbError("rope needs float32 cos / sin tables")
        _contig(cos, sin)
        if rope_n is None:
            rope_n = cos.shape[0]
        width = D // 2 if rope == 2 else D
        if tuple(cos.shape) != (rope_n, width) or tuple(sin.shape) != (rope_n, width):
            raise SvgbError(f"cos / sin must be [{rope_n}, {width}]")
    else:
        rope_n = 0
    for g, n in ((gamma_q, "gamma_q"), (gamma_k, "gamma_k"), (beta_q, "beta_q"), (beta_k, "beta_k")):
        if g is not None:
            want = HD if norm == NORM_RMS_HIDDEN else D
            if g.dtype != q_in.dtype or g.numel() != want or not g.is_contiguous():
                raise SvgbError(f"{n} must be contiguous [{want}] of the input dtype")
    check(lib().svgb_qkv_prep(q_in.data_ptr(), k_in.data_ptr(), v_in.data_ptr(), q_in.stride(1), q_in.stride(0),
                              out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), S_out * D, heads * S_out * D,
                              B, S_in, heads, D, int(out_row0), int(norm), _p(gamma_q), _p(gamma_k), _p(beta_q),
                              _p(beta_k), float(eps), int(rope), _p(cos), _p(sin), int(rope_lo), int(rope_n),
                              _dt(q_in), _stream(q_in)), "svgb_qkv_prep")
    _bump()
    return out


# ---- Wan transformer-block glue (svgb_layernorm_modulate / svgb_rmsnorm_hidden / svgb_modulate_shift / svgb_gate_residual)
_TORCH_DT = {torch.bfloat16: 0, torch.float16: 1, torch.float32: 3}


def _mod_vec(v, N):
    """scale / shift / gate: float32 [N], [nb, N] or [nb, 1, N] -> (tensor [nb, N], nb)."""
    if v.dtype != torch.float32:
        raise SvgbError("modulation vectors must be float32")
    v = v.reshape(-1, N)
    _contig(v)
    return v, v.shape[0]


def _rows_per_batch(rows, nb):
    if nb == 1:
        return 0
    if rows % nb:
        raise SvgbError("rows not divisible by the number of modulation vectors")
    return rows // nb


def layernorm_modulate(x, weight=None, bias=None, eps=1e-6, scale=None, shift=None, out_dtype=None):
    _need_cuda(x, weight, bias, scale, shift)
    _contig(x, weight, bias)
    N = x.shape[-1]
    rows = x.numel() // N
    rpb = 0
    if scale is not None:
        scale, nb = _mod_vec(scale, N)
        shift, nb2 = _mod_vec(shift, N)
        if nb != nb2:
            raise SvgbError("scale and shift must have the same batch")
        rpb = _rows_per_batch(rows, nb)
    y = torch.empty(x.shape, dtype=out_dtype or torch.float32, device=x.device)
    check(lib().svgb_layernorm_modulate(x.data_ptr(), _dt3(x), _p(weight), _p(bias),
                                        _dt3(weight) if weight is not None else 0, float(eps), _p(scale), _p(shift),
                                        rpb, y.data_ptr(), _TORCH_DT[y.dtype], rows, N, _stream(x)),
          "svgb_layernorm_modulate")
    _bump()
    return y


def rmsnorm_hidden(x, weight, eps, out_dtype=None):
    _need_cuda(x, weight)
    _contig(x, weight)
    N = x.shape[-1]
    y = torch.empty(x.shape, dtype=out_dtype or x.dtype, device=x.device)
    check(lib().svgb_rmsnorm_hidden(x.data_ptr(), _dt3(x), weight.data_ptr(), _dt3(weight), float(eps), y.data_ptr(),
                                    _TORCH_DT[y.dtype], x.numel() // N, N, _stream(x)), "svgb_rmsnorm_hidden")
    _bump()
    return y


def modulate_shift(x, scale, shift, out_dtype=torch.float32):
    _need_cuda(x, scale, shift)
    _contig(x)
    N = x.shape[-1]
    rows = x.numel() // N
    scale, nb = _mod_vec(scale, N)
    shift, _ = _mod_vec(shift, N)
    y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    check(lib().svgb_modulate_shift(x.data_ptr(), _dt3(x), scale.data_ptr(), shift.data_ptr(), _rows_per_batch(rows, nb),
                                    y.data_ptr(), _TORCH_DT[y.dtype], rows, N, _stream(x)), "svgb_modulate_shift")
    _bump()
    return y


def gate_residual(residual, x, gate, out_dtype=torch.float32):
    _need_cuda(residual, x, gate)
    _contig(residual, x)
    if residual.shape != x.shape:
        raise SvgbError("residual and x must have the same shape")
    N = x.shape[-1]
    rows = x.numel() // N
    gate, nb = _mod_vec(gate, N)
    y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    check(lib().svgb_gate_residual(residual.data_ptr(), _dt3(residual), x.data_ptr(), _dt3(x), gate.data_ptr(),
                                   _rows_per_batch(rows, nb), y.data_ptr(), _TORCH_DT[y.dtype], rows, N, _stream(x)),
          "svgb_gate_residual")
    _bump()
    return y
