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
                  ws: Optional[torch.Tensor] = None) -> AttnPlan:
    """block_map [BH,QC,KC] bool/uint8, row_sz [BH,QC], col_sz [BH,KC] (any int dtype) on GPU."""
    _need_cuda(block_map, row_sz, col_sz)
    BH, QC, KC = block_map.shape
    m = block_map.contiguous()
    m = m.view(torch.uint8) if m.dtype == torch.bool else m.to(torch.uint8)
    r = row_sz.reshape(BH, QC).to(torch.int32).contiguous()
    c = col_sz.reshape(BH, KC).to(torch.int32).contiguous()
    nbytes = C.c_size_t()
    check(lib().svgb_attn_plan_varblock_bytes(BH, S, QC, KC, C.byref(nbytes)), "plan_varblock_bytes")
    if ws is None:
        ws = workspace(("vb", BH, S, QC, KC), nbytes.value, m.device)
    desc = Plan()
    check(lib().svgb_attn_plan_varblock(m.data_ptr(), r.data_ptr(), c.data_ptr(), BH, S, QC, KC,
                                        ws.data_ptr(), ws.numel(), C.byref(desc), _stream(m)),
          "svgb_attn_plan_varblock")
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
    check(lib().svgb_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), _p(lse), _p(o_rows),
                              _dt(q), BH, S, D, rs, hs, rs, hs, scale, C.byref(plan.desc),
                              plan.ws.data_ptr(), _stream(q)), "svgb_attn_fwd")
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
