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


class CogSVG1Core(SVG1Core):
    """CogVideoX_SparseAttn_Processor2_0.attention_core_logic (svg/models/cog/attention.py:164-196).

    Differences from the HunyuanVideo / Wan cores, all taken from the reference:
      * text tokens come FIRST (placement kernels: cog/placement.py:35-128; profiling masks: cog/utils.py:61-88);
      * sampled rows are drawn from the whole sequence (`torch.randint(0, seq_len)`, cog/attention.py:124), not from
        the first `sample_mse_max_row` rows;
      * the temporal profiling mask has EMPTY text rows (cog/utils.py:76-86): whenever a text row is sampled the
        reference's masked softmax is NaN for every head, the temporal MSE is NaN and `torch.argmin` returns the
        NaN's index -> every head is placed as temporal.  Reproduced here on the host (the kernel itself returns 0
        for rows without keys);
      * the dense switch scales the thresholds: `layer_idx < 42 * first_layers_fp`,
        `timestep[0] > 1000 * (1 - first_times_fp)` (cog/attention.py:172-175), dense = plain SDPA."""
    text_first = True
    smse_layout = 2

    def __init__(self, context_length, num_frame, frame_size, num_heads, head_dim, sparsity, device, cfg_size=1,
                 dtype=torch.bfloat16, num_sampled_rows=32, **kw):
        super().__init__(context_length, num_frame, frame_size, num_sampled_rows=num_sampled_rows,
                         sample_mse_max_row=context_length + num_frame * frame_size, **kw)
        w = sparsity_to_width(sparsity, context_length, num_frame, frame_size)
        self.block_mask = prepare_flexattention(cfg_size, num_heads, head_dim, dtype, device, context_length, num_frame,
                                                frame_size, diag_width=w, multiplier=w)

    def sample_mse(self, query, key, value, sampled_rows=None):
        cfg, H, S, D = query.shape
        if sampled_rows is None:
            sampled_rows = torch.randint(low=0, high=S, size=(min(self.num_sampled_rows, S),))
        mses = super().sample_mse(query, key, value, sampled_rows)
        # the reference's all-masked softmax rows (see the class docstring); no host sync
        text_row = (sampled_rows.to(mses.device) < self.context_length).any()
        mses[1] = torch.where(text_row, torch.full_like(mses[1], float("nan")), mses[1])
        return mses

    def attention_core_logic(self, query, key, value, timestep, layer_idx=None, cu_max_seqlens=None):
        cfg, H, S, D = query.shape
        assert S == self.context_length + self.num_frame * self.frame_size, (
            f"Query Shape: {S} is not equivalent to {self.context_length} + {self.num_frame} * {self.frame_size}")
        full = self.layer_idx < 42 * self.first_layers_fp or bool(timestep[0] > 1000 * (1 - self.first_times_fp))
        if full:
            return self.dense_core(query, key, value, None)
        return self.sparse_core(query, key, value)
