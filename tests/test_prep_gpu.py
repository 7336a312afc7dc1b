"""GPU parity for the pre-attention chain and the Wan block glue.  Parameter grids follow
the reference's own tests (svg/kernels/test/test_{rms_norm,layer_norm,apply_rope,apply_rope_txtlast,
apply_rope_complex}.py) with their tolerances: fp16 (5e-3, 5e-3), bf16 (3e-2, 2e-2)."""
from itertools import product

import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = {torch.float16: (5e-3, 5e-3), torch.bfloat16: (3e-2, 2e-2), torch.float32: (1e-3, 1e-3)}


def close(a, b, dtype=None):
    rtol, atol = TOL[dtype or a.dtype]
    torch.testing.assert_close(a.float().cpu(), b.float().cpu(), rtol=rtol, atol=atol)


def mostly_equal(a, b, frac=0.02):
    """the kernel and the oracle round the same fp32 value once; rsqrt.approx / summation order flip a few ulps"""
    assert (a.float().cpu() != b.float().cpu()).float().mean().item() <= frac


@pytest.mark.parametrize("m, n", list(product([1, 7, 31, 55, 95, 128, 512, 4099], [32, 64, 128, 256])))
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_rms_and_layer_norm(cuda, m, n, dtype):
    from oracle import prep as op
    from svgb200 import _kernels

    g = torch.Generator().manual_seed(m * 1000 + n)
    x = torch.randn(m, n, generator=g).to(dtype)
    gm, bt = torch.randn(n, generator=g).to(dtype), torch.randn(n, generator=g).to(dtype)
    xr = x.clone().to(cuda)
    _kernels.rms_norm_forward(xr, gm.to(cuda), 1e-5)
    close(xr, torch.nn.functional.rms_norm(x.float(), [n], gm.float(), 1e-5).to(dtype))   # test_rms_norm.py:26
    close(xr, op.rms_norm(x, gm, 1e-5))
    mostly_equal(xr, op.rms_norm(x, gm, 1e-5))
    xl = x.clone().to(cuda)
    _kernels.layer_norm_forward(xl, gm.to(cuda), bt.to(cuda))
    close(xl, torch.nn.functional.layer_norm(x, [n], gm, bt, 1e-5))                          # test_layer_norm.py:23
    mostly_equal(xl, op.layer_norm(x, gm, bt), 0.05)


ROPE_GRID = list(product([1, 3], [16], [151, 1037, 6778], [64, 128, 256], [15, 77]))


@pytest.mark.parametrize("bsz, heads, S, D, txt", ROPE_GRID)
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_qk_rope_inplace(cuda, bsz, heads, S, D, txt, mode):
    from oracle import prep as op
    from svgb200 import _kernels

    dtype = torch.float16 if mode == 2 else torch.bfloat16   # as the reference tests choose
    g = torch.Generator().manual_seed(S + D + txt)
    q = torch.randn(bsz, heads, S, D, generator=g).to(dtype)
    k = torch.randn(bsz, heads // 2, S, D, generator=g).to(dtype)   # fewer kv heads exercises Hq != Hk
    width = D // 2 if mode == 2 else D
    cos, sin = torch.randn(S - txt, width, generator=g), torch.randn(S - txt, width, generator=g)
    qo, ko = op.qk_rope_inplace(q, k, cos, sin, txt, mode)
    qd, kd = q.to(cuda), k.to(cuda)
    fn = [_kernels.apply_qk_rope_inplace_cossin, _kernels.apply_qk_rope_inplace_cossin_txtlast,
          _kernels.apply_qk_rope_inplace_cossin_complex][mode]
    fn(qd, kd, cos.to(cuda), sin.to(cuda), txt)
    close(qd, qo)
    close(kd, ko)
    mostly_equal(qd, qo, 0.01)
    text = slice(S - txt, S) if mode == 1 else slice(0, txt)
    assert torch.equal(qd[:, :, text].cpu(), q[:, :, text]) and torch.equal(kd[:, :, text].cpu(), k[:, :, text])


def test_unsupported_head_dim_raises(cuda):
    from svgb200 import _kernels
    from svgb200._lib import SvgbError

    x = torch.zeros(4, 48, device=cuda, dtype=torch.bfloat16)
    with pytest.raises(SvgbError, match="Unsupported head_dim"):
        _kernels.rms_norm_forward(x, torch.zeros(48, device=cuda, dtype=torch.bfloat16), 1e-5)
    with pytest.raises(SvgbError):
        _kernels.rms_norm_forward(x.cpu(), torch.zeros(48, dtype=torch.bfloat16), 1e-5)   # no CPU fallback


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("D, norm, rope", [(128, 1, 1), (64, 2, 1), (128, 3, 2), (128, 0, 0), (256, 1, 2), (64, 0, 1)])
def test_fused_prep_equals_stepwise(cuda, dtype, D, norm, rope):
    """One pass == transpose -> in-place norm -> in-place RoPE, bit for bit (the full-hidden RMS variant is
    compared with the oracle instead: its reduction order differs from the standalone glue kernel)."""
    from oracle import prep as op
    from svgb200 import _kernels, core

    B, S, H, txt = 2, 333, 5, 40
    g = torch.Generator().manual_seed(D + norm * 10 + rope)
    packed = torch.randn(B, S, 3 * H * D, generator=g).to(dtype).to(cuda)      # packed QKV projection output
    qi, ki, vi = packed[..., :H * D], packed[..., H * D:2 * H * D], packed[..., 2 * H * D:]
    gw = H * D if norm == 3 else D
    gq, gk, bq, bk = (torch.randn(gw, generator=g).to(dtype).to(cuda) for _ in range(4))
    width = D // 2 if rope == 2 else D
    cos, sin = (torch.randn(S - txt, width, generator=g).to(cuda) for _ in range(2))
    lo = txt if rope == 2 else 0                                               # complex: text first; else text last
    q, k, v = core.qkv_prep(qi, ki, vi, H, norm=norm, gamma_q=gq if norm else None, gamma_k=gk if norm else None,
                            beta_q=bq if norm == 2 else None, beta_k=bk if norm == 2 else None, eps=1e-6,
                            rope=rope, cos=cos if rope else None, sin=sin if rope else None, rope_lo=lo,
                            rope_n=S - txt if rope else None)
    oq, ok_, ov = op.qkv_chain(qi.cpu(), ki.cpu(), vi.cpu(), H, norm, gq.cpu(), gk.cpu(), bq.cpu(), bk.cpu(), 1e-6,
                               rope, cos.cpu(), sin.cpu(), lo, S - txt)
    close(q, oq), close(k, ok_)
    assert torch.equal(v.cpu(), ov)
    if norm == 3:
        return
    sq, sk = (t.unflatten(2, (H, -1)).transpose(1, 2).contiguous() for t in (qi, ki))
    if norm == 1:
        _kernels.rms_norm_forward(sq.view(-1, D), gq, 1e-6), _kernels.rms_norm_forward(sk.view(-1, D), gk, 1e-6)
    elif norm == 2:
        _kernels.layer_norm_forward(sq.view(-1, D), gq, bq), _kernels.layer_norm_forward(sk.view(-1, D), gk, bk)
    if rope == 1:
        _kernels.apply_qk_rope_inplace_cossin_txtlast(sq, sk, cos, sin, txt)
    elif rope == 2:
        _kernels.apply_qk_rope_inplace_cossin_complex(sq, sk, cos, sin, txt)
    assert torch.equal(q, sq) and torch.equal(k, sk)


def test_double_block_streams_land_in_one_tensor(cuda):
    from svgb200 import _kernels, prep

    B, Sv, St, H, D = 1, 500, 77, 4, 128
    g = torch.Generator(device=cuda).manual_seed(3)
    vid = [torch.randn(B, Sv, H * D, device=cuda, generator=g).bfloat16() for _ in range(3)]
    txt = [torch.randn(B, St, H * D, device=cuda, generator=g).bfloat16() for _ in range(3)]
    w = [torch.randn(D, device=cuda, generator=g).bfloat16() for _ in range(4)]
    cos, sin = torch.randn(Sv, D, device=cuda, generator=g), torch.randn(Sv, D, device=cuda, generator=g)
    q, k, v = prep.hunyuan_double_block_qkv(*vid, *txt, H, w[0], w[1], w[2], w[3], 1e-6, cos, sin)
    # reference sequence (hyvideo/attention.py:253-301)
    sq, sk, sv = (t.unflatten(2, (H, -1)).transpose(1, 2).contiguous() for t in vid)
    _kernels.rms_norm_forward(sq.view(-1, D), w[0], 1e-6), _kernels.rms_norm_forward(sk.view(-1, D), w[1], 1e-6)
    _kernels.apply_qk_rope_inplace_cossin_txtlast(sq, sk, cos, sin, 0)
    eq, ek, ev = (t.unflatten(2, (H, -1)).transpose(1, 2).contiguous() for t in txt)
    _kernels.rms_norm_forward(eq.view(-1, D), w[2], 1e-6), _kernels.rms_norm_forward(ek.view(-1, D), w[3], 1e-6)
    assert torch.equal(q, torch.cat([sq, eq], 2)) and torch.equal(k, torch.cat([sk, ek], 2))
    assert torch.equal(v, torch.cat([sv, ev], 2))


def test_fused_prep_fullsize_hunyuan(cuda):
    """HunyuanVideo 720p single-stream block: S = 119056, H = 24, D = 128 (BASELINE config 2)."""
    from svgb200 import _kernels, prep

    S, H, D, txt = 119056, 24, 128, 256
    g = torch.Generator(device=cuda).manual_seed(4)
    qi, ki, vi = (torch.randn(1, S, H * D, device=cuda, generator=g).bfloat16() for _ in range(3))
    gq, gk = (torch.randn(D, device=cuda, generator=g).bfloat16() for _ in range(2))
    cos, sin = (torch.randn(S - txt, D, device=cuda, generator=g) for _ in range(2))
    q, k, v = prep.hunyuan_single_block_qkv(qi, ki, vi, H, gq, gk, 1e-6, cos, sin, txt)
    assert torch.equal(v, vi.unflatten(2, (H, -1)).transpose(1, 2))
    del v, vi
    for fused, src, gm in ((q, qi, gq), (k, ki, gk)):
        s = src.unflatten(2, (H, -1)).transpose(1, 2).contiguous()
        _kernels.rms_norm_forward(s.view(-1, D), gm, 1e-6)
        _kernels.apply_qk_rope_inplace_cossin_txtlast(s, s[:, :1].clone(), cos, sin, txt)
        assert torch.equal(fused, s)
        del s


@pytest.mark.parametrize("N", [1536, 5120, 8192, 264])
@pytest.mark.parametrize("x_dtype", [torch.bfloat16, torch.float32])
def test_wan_block_glue(cuda, N, x_dtype):
    from oracle import prep as op
    from svgb200 import triton_glue as tg

    B, S = 2, 37
    g = torch.Generator().manual_seed(N)
    x = torch.randn(B, S, N, generator=g).to(x_dtype)
    r = torch.randn(B, S, N, generator=g).bfloat16()
    w, b = torch.randn(N, generator=g), torch.randn(N, generator=g)
    scale, shift, gate = (torch.randn(B, 1, N, generator=g) for _ in range(3))
    xd = x.to(cuda)
    close(tg.triton_layernorm_forward(xd, None, None, 1e-6, False), op.layernorm_hidden(x, None, None, 1e-6))
    close(tg.triton_layernorm_forward(xd, w.to(cuda), b.to(cuda), 1e-6, True), op.layernorm_hidden(x, w, b, 1e-6))
    n32 = op.layernorm_hidden(x, None, None, 1e-6)
    close(tg.triton_modulate_shift_forward(n32.to(cuda), scale.to(cuda), shift.to(cuda), torch.bfloat16),
          op.modulate_shift(n32, scale, shift, torch.bfloat16))
    close(tg.layernorm_modulate_forward(xd, None, None, 1e-6, scale.to(cuda), shift.to(cuda), torch.bfloat16),
          op.modulate_shift(n32, scale, shift, torch.bfloat16))
    close(tg.triton_modulate_gate_residual_forward(r.to(cuda), xd, gate.to(cuda), torch.bfloat16),
          op.gate_residual(r, x, gate, torch.bfloat16))
    close(tg.triton_rmsnorm_forward(xd, w.to(x_dtype).to(cuda), 1e-6), op.rmsnorm_hidden(x, w.to(x_dtype), 1e-6))
    # single modulation vector for every row (the reference kernels' own indexing, modulate.py:30-31)
    close(tg.triton_modulate_shift_forward(xd, scale[0, 0].to(cuda), shift[0, 0].to(cuda), torch.float32),
          op.modulate_shift(x, scale[0, 0], shift[0, 0]))
