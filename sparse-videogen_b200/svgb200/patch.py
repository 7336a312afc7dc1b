"""Install the svgb200 operators into an importable copy of the reference (svg-project/Sparse-VideoGen).

The reference's attention processors resolve their operators as module globals
(svg/models/hyvideo/attention.py:12-28); replacing those names is the same mechanism the reference uses
for its own monkey patches (custom_models.py:259-263).  No reference file is edited.

    import svg.models.hyvideo.attention as A
    import svgb200.patch as P
    P.install(A)            # hyvideo / wan / cosmos / cog attention modules
    P.install_kmeans_utils()  # svg.kmeans_utils names used elsewhere
"""
from __future__ import annotations

from . import kmeans_utils as _ku
from . import permute as _pm
from . import placement as _pl

_NAMES = {
    "apply_inverse_permutation_triton": _pm.apply_inverse_permutation_triton,
    "permute_tensor_by_labels_triton": _pm.permute_tensor_by_labels_triton,
    "batch_kmeans_Euclid": _ku.batch_kmeans_Euclid,
    "density_calculation": _ku.density_calculation,
    "dynamic_block_sparse_fwd_flashinfer": _ku.dynamic_block_sparse_fwd_flashinfer,
    "dynamic_block_sparse_fwd_triton": _ku.dynamic_block_sparse_fwd_triton,
    "identify_dynamic_map": _ku.identify_dynamic_map,
    "hunyuan_sparse_head_placement": _pl.hunyuan_sparse_head_placement,
    "hunyuan_hidden_states_placement": _pl.hunyuan_hidden_states_placement,
    "wan_sparse_head_placement": _pl.wan_sparse_head_placement,
    "wan_hidden_states_placement": _pl.wan_hidden_states_placement,
    "cosmos_sparse_head_placement": _pl.cosmos_sparse_head_placement,
    "cosmos_hidden_states_placement": _pl.cosmos_hidden_states_placement,
    "sparse_head_placement": _pl.sparse_head_placement,
    "hidden_states_placement": _pl.hidden_states_placement,
}


def flex_attention(query, key, value, block_mask=None, **kw):
    """Stand-in for torch.nn.attention.flex_attention.flex_attention as the reference processors call it
    (`flex_attention(q, k, v, block_mask=cls.block_mask)`, hyvideo/attention.py:401-403): block_mask must be the
    BandMask built by the patched prepare_flexattention."""
    from .models.common import BandMask, sparse_flex_attention

    if not isinstance(block_mask, BandMask):
        raise TypeError("svgb200.patch.flex_attention needs the BandMask returned by the patched prepare_flexattention "
                        f"(got {type(block_mask).__name__}); rebuild block_mask after patch.install()")
    if kw:
        raise TypeError(f"unsupported flex_attention arguments: {sorted(kw)}")
    return sparse_flex_attention(query, key, value, block_mask)


def flashinfer_varlen_func(q, k, v, cu_seqlens_q, cu_seqlens_kv, max_seqlen_q, max_seqlen_kv):
    """hyvideo/attention.py:807-875: dense attention inside each cu_seqlens segment, q,k,v [B,H,S,D]."""
    from .models.common import dense_attention

    seg = [n for n in (cu_seqlens_q[1:] - cu_seqlens_q[:-1]).tolist() if n > 0]
    return dense_attention(q, k, v, seg)


def flash_attn_varlen_func(q, k, v, cu_seqlens_q, cu_seqlens_kv, max_seqlen_q, max_seqlen_kv, **kw):
    """flash_attn.flash_attn_interface.flash_attn_varlen_func as hyvideo/attention.py:452-470 calls it: q,k,v
    [L, B*H, D] (sequence first), segments given by cu_seqlens."""
    from . import core

    S, BH, D = q.shape
    seg = [n for n in (cu_seqlens_q[1:] - cu_seqlens_q[:-1]).tolist() if n > 0]
    dev = q.device
    n = len(seg)
    import torch

    bm = torch.eye(n, dtype=torch.bool, device=dev).expand(BH, n, n).contiguous()
    sz = torch.tensor(seg, dtype=torch.int32, device=dev).expand(BH, n).contiguous()
    plan = core.plan_varblock(bm, sz, sz, S)
    return core.attn_fwd(q.contiguous(), k.contiguous(), v.contiguous(), plan, layout="shd")


def _model_of(module):
    name = getattr(module, "__name__", "")
    for m in ("hyvideo", "wan", "cosmos", "cog"):
        if f".{m}." in name or name.endswith("." + m):
            return m
    return None


def install(module, sample_mse: bool = True) -> list:
    """Replace every operator name the module already defines (the reference resolves them as module globals,
    svg/models/hyvideo/attention.py:12-28).  Besides the kernels this covers the module-level attention entry points
    `flex_attention`, `prepare_flexattention`, `flashinfer_varlen_func`, `flash_attn_varlen_func`, and — with
    sample_mse=True — the processors' `sample_mse` method (the reference version needs the materialised
    [10000, S] profiling masks in `attention_masks`; ours evaluates them analytically).  Returns the names replaced."""
    done = []
    for name, fn in _NAMES.items():
        if hasattr(module, name):
            setattr(module, name, fn)
            done.append(name)
    for name, fn in (("flex_attention", flex_attention), ("flashinfer_varlen_func", flashinfer_varlen_func),
                     ("flash_attn_varlen_func", flash_attn_varlen_func)):
        if hasattr(module, name):
            setattr(module, name, fn)
            done.append(name)
    model = _model_of(module)
    if model is not None and hasattr(module, "prepare_flexattention"):
        from .models import cog, hyvideo, wan

        module.prepare_flexattention = {"hyvideo": hyvideo, "wan": wan, "cosmos": wan, "cog": cog}[model].prepare_flexattention
        done.append("prepare_flexattention")
    if sample_mse and model is not None:
        layout = {"hyvideo": 0, "wan": 1, "cosmos": 1, "cog": 2}[model]
        for obj in vars(module).values():
            if isinstance(obj, type) and "sample_mse" in vars(obj):
                obj.sample_mse = _make_sample_mse(layout)
                done.append(f"{obj.__name__}.sample_mse")
    return done


def _make_sample_mse(layout: int):
    def sample_mse(self, query, key, value):
        """<Model>_SVGAttn_Processor.sample_mse (hyvideo/attention.py:375-399) on the svgb200 kernel: same sampled
        rows (CPU generator, :381), same [2, cfg, H] result in query.dtype; the profiling masks are evaluated
        analytically, `attention_masks` is not needed."""
        import torch

        from . import core

        cfg, H, S, D = query.shape
        n = min(self.num_sampled_rows, S)
        high = S if layout == 2 else self.sample_mse_max_row  # cog/attention.py:124 samples the whole sequence
        rows = torch.randint(low=0, high=high, size=(n,))
        mse = core.sample_mse(query.reshape(cfg * H, S, D), key.reshape(cfg * H, S, D), value.reshape(cfg * H, S, D),
                              rows.to(query.device), layout, self.context_length, self.num_frame, self.frame_size)
        mse = mse.view(2, cfg, H).to(query.dtype)
        if layout == 2:  # empty text rows of the Cog temporal mask -> NaN in the reference (cog/utils.py:76-86)
            text_row = (rows < self.context_length).any().to(query.device)
            mse[1] = torch.where(text_row, torch.full_like(mse[1], float("nan")), mse[1])
        return mse

    return sample_mse


def install_kmeans_utils(module=None) -> list:
    if module is None:
        import svg.kmeans_utils as module  # noqa: WPS433 (the reference must be importable)
    names = ["batch_kmeans_Euclid", "density_calculation", "dynamic_block_sparse_fwd_flashinfer",
             "dynamic_block_sparse_fwd_triton", "identify_dynamic_map", "euclid_assign_triton",
             "triton_centroid_update_sorted_euclid"]
    for n in names:
        setattr(module, n, getattr(_ku, n))
    return names


def install_native_kernels() -> None:
    """Make `import _kernels` (svg/models/*/attention.py: `sys.path.append("svg/kernels/build/"); import _kernels`)
    resolve to the svgb200 implementation, so the reference's ENABLE_FAST_KERNEL branch runs these kernels."""
    import sys

    from . import _kernels as k

    sys.modules["_kernels"] = k


def install_triton_glue(module) -> list:
    """Replace the Triton glue names a reference module imported (svg/models/wan/custom_models.py:16-17,
    svg/models/wan/attention.py:16)."""
    from . import triton_glue as g

    done = []
    for name in ("triton_rmsnorm_forward", "triton_layernorm_forward", "triton_modulate_shift_forward",
                 "triton_modulate_gate_residual_forward"):
        if hasattr(module, name):
            setattr(module, name, getattr(g, name))
            done.append(name)
    return done
