"""Drop-in for the reference's native pybind module `_kernels` (svg/kernels/csrc/ops.cu:3-9, ops.h): same
function names, argument order and in-place semantics, backed by libsvgb200.so.

    import svgb200._kernels as _kernels          # or svgb200.patch.install_native_kernels()
    _kernels.rms_norm_forward(x.view(-1, D), weight, eps)
    _kernels.apply_qk_rope_inplace_cossin_txtlast(q, k, cos, sin, txt_len)
"""
from __future__ import annotations

import torch

from . import core
from ._lib import SvgbError


def _check_2d(input, *vecs):
    if input.dim() != 2:
        raise SvgbError("input must be [m, n]")
    for v in vecs:
        if v.dim() != 1 or v.shape[0] != input.shape[1]:
            raise SvgbError("gamma / beta must be [n]")


def rms_norm_forward(input: torch.Tensor, gemma: torch.Tensor, epsilon: float = 1e-5) -> None:
    """ops.h:52-75 — in place on input [m, n]."""
    _check_2d(input, gemma)
    core.rms_norm_(input, gemma, epsilon)


def layer_norm_forward(input: torch.Tensor, gemma: torch.Tensor, beta: torch.Tensor) -> None:
    """ops.h:20-44 — in place on input [m, n], eps = 1e-5."""
    _check_2d(input, gemma, beta)
    core.layer_norm_(input, gemma, beta)


def apply_qk_rope_inplace_cossin(q, k, cos_cache, sin_cache, len_text_prompt: int) -> None:
    """ops.h:77-133 — the FIRST len_text_prompt rows of every head are skipped."""
    core.qk_rope_(q, k, cos_cache, sin_cache, len_text_prompt, core.ROPE_TXT_FIRST)


def apply_qk_rope_inplace_cossin_txtlast(q, k, cos_cache, sin_cache, len_text_prompt: int) -> None:
    """ops.h:135-197 — the LAST len_text_prompt rows of every head are skipped."""
    core.qk_rope_(q, k, cos_cache, sin_cache, len_text_prompt, core.ROPE_TXT_LAST)


def apply_qk_rope_inplace_cossin_complex(q, k, cos_cache, sin_cache, len_text_prompt: int) -> None:
    """ops.h:199-260 — cos/sin are the real / imaginary parts [valid, D/2]; fp64 arithmetic."""
    core.qk_rope_(q, k, cos_cache, sin_cache, len_text_prompt, core.ROPE_COMPLEX_TXT_FIRST)
