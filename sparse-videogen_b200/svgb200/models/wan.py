"""Wan 2.1 / Cosmos (no text tokens in self-attention, first-frame sink) — mirrors
svg/models/{wan,cosmos}/{attention,utils,inference}.py."""
from __future__ import annotations

from math import ceil

import torch

from .. import core
from ..placement import wan_hidden_states_placement, wan_sparse_head_placement  # noqa: F401
from .common import BandMask, KMeansState, SAPCore, SVG1Core, sparse_flex_attention, sparsity_to_width  # noqa: F401


def gen_temporal_mask(num_frames: int, num_tokens_per_frame: int, multiplier: float, device=None):
    """svg/models/wan/utils.py:130-185: BSR temporal mask = diagonal band + first-frame region."""
    from ..ops.attention_ops_wan import gen_temporal_mask as _g

    return _g(num_frames, num_tokens_per_frame, multiplier, device, first_frame=True)


from ..ops.attention_ops_wan import flashinfer_sparse_attn_forward  # noqa: E402,F401  (wan/utils.py:188-238)


def band_params(num_frames, token_per_frame, mul):
    """generate_temporal_head_mask_mod (wan/utils.py:25-41): kv < P | |q-kv| <= ceil(mul*P/128)*128."""
    return core.MASK_WAN, token_per_frame, 0, ceil(mul * token_per_frame / 128) * 128


def prepare_flexattention(cfg_size, num_head, head_dim, dtype, device, context_length, prompt_length, num_frame,
                          frame_size, diag_width=1, multiplier=2) -> BandMask:
    """wan/attention.py prepare_flexattention."""
    assert diag_width == multiplier
    S = context_length + num_frame * frame_size
    mode, m0, m1, m2 = band_params(num_frame, frame_size, multiplier)
    plan = core.plan_band(mode, m0, m1, m2, cfg_size * num_head, S, device)
    return BandMask(plan, mode, m0, m1, m2, S)


class WanSVG1Core(SVG1Core):
    """Wan_SVGAttn_Processor2_0.attention_core_logic (wan/attention.py:284-328)."""
    text_first = False
    smse_layout = 1

    def __init__(self, num_frame, frame_size, num_heads, head_dim, sparsity, device, cfg_size=1,
                 dtype=torch.bfloat16, **kw):
        super().__init__(0, num_frame, frame_size, **kw)
        w = sparsity_to_width(sparsity, 0, num_frame, frame_size)
        self.block_mask = prepare_flexattention(cfg_size, num_heads, head_dim, dtype, device, 0, 0, num_frame,
                                                frame_size, diag_width=w, multiplier=w)


class WanSAPCore(SAPCore):
    """Wan_SAPAttn_Processor2_0.attention_core_logic (wan/attention.py:499-559); centroids per instance."""

    def __init__(self, num_frame, frame_size, **kw):
        super().__init__(0, num_frame, frame_size, **kw)


CosmosSVG1Core = WanSVG1Core
CosmosSAPCore = WanSAPCore
