"""Mirror of the reference's Triton block-glue kernels (svg/kernels/triton/{rmsnorm,layernorm,modulate}.py), used
by WanTransformerBlock_Sparse.forward (svg/models/wan/custom_models.py:37-111): same names and arguments, CUDA
kernels from libsvgb200.so.  `layernorm_modulate_forward` is the B200 form of the pair the block always calls
back to back (LayerNorm -> x*(1+scale)+shift): one pass, no fp32 intermediate in HBM.
"""
from __future__ import annotations

import torch

from . import core


def triton_rmsnorm_forward(x, w, eps):
    """rmsnorm.py:52-103 — RMS over the last (full hidden) dim; output in x's dtype."""
    return core.rmsnorm_hidden(x, w, eps)


def triton_layernorm_param_forward(x, w, b, eps):
    """layernorm.py:64-104 — float32 result."""
    return core.layernorm_modulate(x, w, b, eps, out_dtype=torch.float32)


def triton_layernorm_noparam_forward(x, eps):
    """layernorm.py:157-197 — float32 result."""
    return core.layernorm_modulate(x, None, None, eps, out_dtype=torch.float32)


def triton_layernorm_forward(x, w, b, eps, elementwise_affine=True):
    """layernorm.py:204-210."""
    if elementwise_affine:
        assert w is not None and b is not None
        return triton_layernorm_param_forward(x, w, b, eps)
    assert w is None and b is None
    return triton_layernorm_noparam_forward(x, eps)


def triton_modulate_shift_forward(x, scale, shift, output_dtype=torch.float32):
    """modulate.py:42-75 — y = x * (1 + scale) + shift."""
    return core.modulate_shift(x, scale, shift, out_dtype=output_dtype)


def triton_modulate_gate_residual_forward(residual, x, gate, output_dtype=torch.float32):
    """modulate.py:117-152 — y = residual + x * gate."""
    return core.gate_residual(residual, x, gate, out_dtype=output_dtype)


def layernorm_modulate_forward(x, w, b, eps, scale, shift, output_dtype=None):
    """custom_models.py:37-56 fused: LayerNorm (optional affine) then x*(1+scale)+shift, rounded once to
    `output_dtype` (default: x's dtype, as `.type_as(hidden_states)` does)."""
    return core.layernorm_modulate(x, w, b, eps, scale, shift, out_dtype=output_dtype or x.dtype)
