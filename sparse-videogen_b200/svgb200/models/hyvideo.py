"""HunyuanVideo (text tokens LAST) — mirrors svg/models/hyvideo/{attention,utils,inference}.py."""
from __future__ import annotations

from math import floor

import torch

from .. import core
from ..placement import hunyuan_hidden_states_placement, hunyuan_sparse_head_placement  # noqa: F401
from .common import BandMask, KMeansState, SAPCore, SVG1Core, sparse_flex_attention, sparsity_to_width  # noqa: F401


def band_params(context_length, prompt_length, num_frames, token_per_frame, mul):
    """generate_temporal_head_mask_mod (hyvideo/utils.py:20-44) as engine parameters."""
    V = num_frames * token_per_frame
    return core.MASK_HY, V, V + int(prompt_length), floor(mul * token_per_frame / 128) * 128


def prepare_flexattention(cfg_size, num_head, head_dim, dtype, device, context_length, prompt_length, num_frame,
                          frame_size, diag_width=1, multiplier=2) -> BandMask:
    """hyvideo/attention.py:527-551: build the executed SVG1 mask once."""
    assert diag_width == multiplier
    S = context_length + num_frame * frame_size
    mode, m0, m1, m2 = band_params(context_length, prompt_length, num_frame, frame_size, multiplier)
    plan = core.plan_band(mode, m0, m1, m2, cfg_size * num_head, S, device)
    return BandMask(plan, mode, m0, m1, m2, S)


class HunyuanSVG1Core(SVG1Core):
    """Hunyuan_SVGAttn_Processor2_0.attention_core_logic (hyvideo/attention.py:473-524)."""
    text_first = False
    smse_layout = 0

    def __init__(self, context_length, prompt_length, num_frame, frame_size, num_heads, head_dim, sparsity,
                 device, cfg_size=1, dtype=torch.bfloat16, **kw):
        super().__init__(context_length, num_frame, frame_size, **kw)
        self.prompt_length = prompt_length
        w = sparsity_to_width(sparsity, context_length, num_frame, frame_size)
        self.block_mask = prepare_flexattention(cfg_size, num_heads, head_dim, dtype, device, context_length,
                                                prompt_length, num_frame, frame_size, diag_width=w, multiplier=w)


class HunyuanSAPCore(SAPCore):
    """Hunyuan_SAPAttn_Processor2_0.attention_core_logic (hyvideo/attention.py:714-804)."""
