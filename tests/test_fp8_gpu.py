"""FP8 (e4m3) attention path (BASELINE config 5).  The reference has no FP8 sparse attention
(README.md:117; FlashInfer's fp8 sparse path asserts backend == "fa3"), so the oracle is fp32 attention
on the DEQUANTISED inputs and the tolerance is ours: P is rounded to e4m3 (3 mantissa bits, relative
step 2^-3..2^-4) before PV, which averages out over the keys of a row; we require rtol 8e-2 / atol 4e-2
on outputs of magnitude ~0.1-1 and report the measured error."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_quantize_e4m3(cuda):
    from svgb200 import core

    g = torch.Generator().manual_seed(0)
    x = (torch.randn(2, 3, 500, 128, generator=g) * torch.tensor([0.1, 1.0, 30.0]).view(1, 3, 1, 1)).bfloat16()
    x8, scale = core.quantize_e4m3(x.to(cuda))
    deq = x8.view(torch.float8_e4m3fn).float().cpu() * scale.cpu().view(2, 3, 1, 1)
    amax = x.float().abs().amax(dim=(2, 3)).view(-1)
    torch.testing.assert_close(scale.cpu(), amax / 448.0, rtol=1e-6, atol=0)
    err = (deq - x.float()).abs()
    bound = torch.maximum(x.float().abs() * 2 ** -4, scale.cpu().view(2, 3, 1, 1) * 2 ** -10)  # half ulp (3 mantissa bits)
    assert bool((err <= bound * 1.001 + 1e-12).all())
    assert float(deq.abs().amax()) == pytest.approx(float(x.float().abs().amax()), rel=1e-6)


@pytest.mark.parametrize("S,MB,NB,density", [(1024, 5, 9, 0.4), (4096, 20, 100, 0.7)])
def test_fp8_variable_block_attention(cuda, S, MB, NB, density):
    from oracle.attention import dynamic_block_sparse_fwd
    from svgb200 import core

    g = torch.Generator().manual_seed(S)
    H, D = 2, 128
    q, k, v = (torch.randn(1, H, S, D, generator=g).bfloat16() for _ in range(3))

    def part(n):
        cuts = torch.sort(torch.randperm(S - 1, generator=g)[: n - 1] + 1)[0]
        return torch.diff(torch.cat([torch.tensor([0]), cuts, torch.tensor([S])])).to(torch.int32)
    row = torch.stack([part(MB) for _ in range(H)])
    col = torch.stack([part(NB) for _ in range(H)])
    bm = torch.rand(H, MB, NB, generator=g) > density
    (q8, sq), (k8, sk), (v8, sv) = (core.quantize_e4m3(t.to(cuda)) for t in (q, k, v))
    plan = core.plan_varblock(bm.to(cuda), row.to(cuda), col.to(cuda), S)
    o = core.attn_fwd_fp8(q8, k8, v8, sq, sk, sv, plan).float().cpu()

    def deq(x8, s):
        return x8.view(torch.float8_e4m3fn).float().cpu() * s.cpu().view(1, H, 1, 1)
    ref = dynamic_block_sparse_fwd(deq(q8, sq), deq(k8, sk), deq(v8, sv), bm[None], row[None], col[None])
    err = (o - ref).abs()
    print("fp8 attention max abs err", float(err.max()), "mean", float(err.mean()))
    torch.testing.assert_close(o, ref, rtol=8e-2, atol=4e-2)
    # and it stays close to the bf16 path on the original inputs
    o16 = core.attn_fwd(q.to(cuda), k.to(cuda), v.to(cuda), plan).float().cpu()
    assert float((o - o16).abs().mean()) < 2e-2


def test_fp8_band_attention(cuda):
    from oracle import attention as oa
    from svgb200 import core
    from svgb200.models import wan

    g = torch.Generator().manual_seed(5)
    H, F, P, D = 2, 6, 200, 128
    S = F * P
    q, k, v = (torch.randn(1, H, S, D, generator=g).bfloat16() for _ in range(3))
    bm = wan.prepare_flexattention(1, H, D, torch.bfloat16, cuda, 0, 0, F, P, diag_width=1.3, multiplier=1.3)
    (q8, sq), (k8, sk), (v8, sv) = (core.quantize_e4m3(t.to(cuda)) for t in (q, k, v))
    o = core.attn_fwd_fp8(q8, k8, v8, sq, sk, sv, bm.plan).float().cpu()

    def deq(x8, s):
        return x8.view(torch.float8_e4m3fn).float().cpu() * s.cpu().view(1, H, 1, 1)
    ref = oa.masked_attention_bhsd(deq(q8, sq)[0], deq(k8, sk)[0], deq(v8, sv)[0], oa.wan_mask_mod(F, P, 1.3))[None]
    torch.testing.assert_close(o, ref, rtol=8e-2, atol=4e-2)
