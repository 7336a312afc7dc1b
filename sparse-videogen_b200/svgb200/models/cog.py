"""CogVideoX (text tokens FIRST, SVG1 only) — mirrors svg/models/cog/{attention,utils,inference}.py."""
from __future__ import annotations

from math import floor

import torch

from .. import core
from ..placement import cog_hidden_states_placement, cog_sparse_head_placement  # noqa: F401
from .common import BandMask, SVG1Core, sparse_flex_attention, sparsity_to_width  # noqa: F401


def band_params(prompt_length, num_frames, token_per_frame, mul, attn_sink=False):
    """generate_temporal_head_mask_mod (cog/utils.py:30-46)."""
    first_col = prompt_length + token_per_frame if attn_sink else prompt_length
    return core.MASK_COG, first_col, prompt_length, floor(mul * token_per_frame / 128) * 128


def prepare_flexattention(cfg_size, num_head, head_dim, dtype, device, context_length, num_frame, frame_size,
                          diag_width=1, multiplier=2, attn_sink=False) -> BandMask:
    assert diag_width == multiplier
    S = context_length + num_frame * frame_size
    mode, m0, m1, m2 = band_params(context_length, num_frame, frame_size, multiplier, attn_sink)
    plan = core.plan_band(mode, m0, m1, m2, cfg_size * num_head, S, device)
    return BandMask(plan, mode, m0, m1, m2, S)
