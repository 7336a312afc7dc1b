"""CPU pinning of oracle/prep.py against golden vectors produced by the reference's own ground-truth functions
(tests/golden/make_golden_prep.py) and against torch's library norms the reference tests compare with."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import prep as op

G = np.load(Path(__file__).resolve().parent / "golden" / "prep_golden.npz")


def T(name, dtype=torch.bfloat16):
    return torch.from_numpy(G[name]).to(dtype)


def close16(a, b, dtype):
    rtol, atol = {torch.float16: (5e-3, 5e-3), torch.bfloat16: (3e-2, 2e-2)}[dtype]  # test_rms_norm.py:14-19
    torch.testing.assert_close(a.float(), b.float(), rtol=rtol, atol=atol)


@pytest.mark.parametrize("n", [32, 64, 128, 256])
def test_norms_match_reference_ground_truth(n):
    x, g, b = T(f"rms_x_{n}"), T(f"rms_g_{n}"), T(f"ln_b_{n}")
    y = op.rms_norm(x, g, 1e-5)
    close16(y, T(f"rms_ref_{n}"), torch.bfloat16)
    # the oracle rounds once (like the CUDA kernel); the reference's "replica" rounds before gamma: <= 1 bf16 ulp apart
    close16(y, T(f"rms_replica_{n}"), torch.bfloat16)
    assert (y.float() != T(f"rms_ref_{n}").float()).float().mean() < 0.02
    close16(op.layer_norm(x, g, b), T(f"ln_ref_{n}"), torch.bfloat16)
    close16(op.layer_norm(x, g, b), torch.nn.functional.layer_norm(x, [n], g, b, 1e-5), torch.bfloat16)


@pytest.mark.parametrize("D", [64, 128])
def test_rope_matches_reference_ground_truth(D):
    q = T(f"rope_q_{D}")
    cos, sin = T(f"rope_cos_{D}", torch.float32), T(f"rope_sin_{D}", torch.float32)
    Tn = q.shape[2] - cos.shape[0]
    last, _ = op.qk_rope_inplace(q, q, cos, sin, Tn, 1)
    assert torch.equal(last[:, :, :-Tn].float(), T(f"rope_txtlast_{D}").float())      # same fp32 formula: bit exact
    assert torch.equal(last[:, :, -Tn:], q[:, :, -Tn:])
    first, _ = op.qk_rope_inplace(q, q, cos, sin, Tn, 0)
    assert torch.equal(first[:, :, Tn:].float(), T(f"rope_txtfirst_{D}").float())
    assert torch.equal(first[:, :, :Tn], q[:, :, :Tn])
    qh = q.half()
    re, im = T(f"ropec_re_{D}", torch.float32), T(f"ropec_im_{D}", torch.float32)
    cplx, _ = op.qk_rope_inplace(qh, qh, re, im, Tn, 2)
    assert torch.equal(cplx[:, :, Tn:].float(), T(f"ropec_out_{D}", torch.float16).float())


def test_chain_equals_the_stepwise_sequence():
    g = torch.Generator().manual_seed(1)
    B, S, H, D, txt = 1, 40, 3, 64, 7
    qi, ki, vi = (torch.randn(B, S, H * D, generator=g).bfloat16() for _ in range(3))
    gq, gk = torch.randn(D, generator=g).bfloat16(), torch.randn(D, generator=g).bfloat16()
    cos, sin = torch.randn(S - txt, D, generator=g), torch.randn(S - txt, D, generator=g)
    q, k, v = op.qkv_chain(qi, ki, vi, H, 1, gq, gk, None, None, 1e-6, 1, cos, sin, 0, S - txt)
    q2 = qi.unflatten(2, (H, -1)).transpose(1, 2).contiguous()
    q2 = op.rms_norm(q2, gq, 1e-6)
    q2, _ = op.qk_rope_inplace(q2, q2, cos, sin, txt, 1)
    assert torch.equal(q, q2)
    assert torch.equal(v, vi.unflatten(2, (H, -1)).transpose(1, 2))


def test_glue_matches_eager_forms():
    """custom_models.py:44-66 (the reference's non-Triton branch) are the eager forms of the glue kernels."""
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 9, 256, generator=g).bfloat16()
    scale, shift, gate = (torch.randn(2, 1, 256, generator=g) for _ in range(3))
    ln = torch.nn.LayerNorm(256, eps=1e-6, elementwise_affine=False)
    n = ln(x.float())
    torch.testing.assert_close(op.layernorm_hidden(x, None, None, 1e-6), n, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(op.modulate_shift(n, scale, shift, torch.bfloat16),
                               (n * (1 + scale) + shift).type_as(x))
    torch.testing.assert_close(op.gate_residual(x, x * 2, gate, torch.bfloat16),
                               (x.float() + (x * 2) * gate).type_as(x))
    w = torch.randn(256, generator=g)
    torch.testing.assert_close(op.rmsnorm_hidden(x.float(), w, 1e-6),
                               torch.nn.functional.rms_norm(x.float(), [256], w, 1e-6), rtol=1e-5, atol=1e-5)
