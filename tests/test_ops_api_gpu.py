"""GPU parity for the mirrored ops API (svg/kernels/ops), following the reference tests
svg/kernels/test/test_sparse_attn.py:164-232 and test_sparse_attn_wan.py:70-100 at reduced size
(the oracle is the same naive torch attention on the block->element expanded mask)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = {torch.float16: dict(rtol=5e-3, atol=5e-3), torch.bfloat16: dict(rtol=3e-2, atol=2e-2)}


@pytest.mark.parametrize("pattern,mul", [("spatial", 2), ("temporal", 1.8)])
@pytest.mark.parametrize("H,D", [(3, 64), (2, 128)])
def test_sparse_attn_forward_text_first(cuda, pattern, mul, H, D):
    from oracle import attention as oa
    from svgb200.ops import FAMetadata, _gen_spatial_mask, _gen_temporal_mask, sparse_attn_forward

    F, P, T = 6, 160, 16
    S = F * P + T
    g = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(S, H, D, generator=g).to(torch.float16) for _ in range(3))
    if pattern == "spatial":
        md = _gen_spatial_mask(F, P, mul, device=cuda)
        ref_bm, bs = oa.ref_gen_spatial_mask(F, P, mul)
        meta = FAMetadata(T, F, P, None, md)
    else:
        md = _gen_temporal_mask(F, P, mul, device=cuda)
        ref_bm, bs = oa.ref_gen_temporal_mask(F, P, mul)
        meta = FAMetadata(T, F, P, md, None)
    assert tuple(md[2]) == tuple(bs)
    o = sparse_attn_forward(q.to(cuda), k.to(cuda), v.to(cuda), meta, pattern).float().cpu()
    ref = oa.ref_torch_attn_impl(q, k, v, oa.gen_mask_block2element(ref_bm, bs, T))
    torch.testing.assert_close(o, ref, **TOL[torch.float16])


@pytest.mark.parametrize("first_frame", [False, True])
@pytest.mark.parametrize("mul", [1.3, 2.2])
def test_wan_sparse_attn_forward(cuda, mul, first_frame):
    """first_frame=False: svg/kernels/ops/attention_ops_wan.gen_temporal_mask (band only);
    True: svg/models/wan/utils.gen_temporal_mask (band + first-frame region)."""
    from oracle import attention as oa
    from svgb200.models import wan as wan_m
    from svgb200.ops import WanFAMetadata, flashinfer_sparse_attn_forward, wan_sparse_attn_forward
    from svgb200.ops import gen_temporal_mask as ops_gen

    gen_temporal_mask = wan_m.gen_temporal_mask if first_frame else ops_gen

    F, P, H, D = 5, 240, 2, 128
    S = F * P
    g = torch.Generator().manual_seed(1)
    q, k, v = (torch.randn(S, H, D, generator=g).to(torch.bfloat16) for _ in range(3))
    md = gen_temporal_mask(F, P, mul, device=cuda)
    ref_bm, bs = oa.ref_gen_temporal_mask_wan(F, P, mul, first_frame=first_frame)
    assert tuple(md[2]) == tuple(bs)
    o = wan_sparse_attn_forward(q.to(cuda), k.to(cuda), v.to(cuda), WanFAMetadata(F, P, md)).float().cpu()
    ref = oa.ref_torch_attn_impl(q, k, v, oa.gen_mask_block2element(ref_bm, bs, 0))
    torch.testing.assert_close(o, ref, **TOL[torch.bfloat16])
    qb = q.permute(1, 0, 2)[None].contiguous()
    o2 = flashinfer_sparse_attn_forward(qb.to(cuda), k.permute(1, 0, 2)[None].contiguous().to(cuda),
                                        v.permute(1, 0, 2)[None].contiguous().to(cuda), md).float().cpu()
    torch.testing.assert_close(o2[0].permute(1, 0, 2), ref, **TOL[torch.bfloat16])


def test_sap_core_wan_pipeline(cuda):
    """Whole SVG2 core (Wan layout, no text): engine output == masked attention under the element mask
    implied by the engine's own integer outputs (labels/sizes/map/permutation)."""
    from svgb200.models import wan

    g = torch.Generator().manual_seed(2)
    H, F, P, D = 3, 4, 300, 128
    S = F * P
    q, k, v = (torch.randn(1, H, S, D, generator=g).to(torch.bfloat16) for _ in range(3))
    sap = wan.WanSAPCore(F, P, num_q_centroids=6, num_k_centroids=20, top_p_kmeans=0.8, min_kc_ratio=0.1,
                         kmeans_iter_init=4, kmeans_iter_step=1)
    for step in range(2):  # second call warm-starts from the stored centroids
        o = sap.sparse_core(q.to(cuda), k.to(cuda), v.to(cuda)).float().cpu()
        qperm, kperm = sap.last["q_sorted_indices"].cpu().long(), sap.last["k_sorted_indices"].cpu().long()
        m, rs, cs = (sap.last[x].cpu() for x in ("dynamic_map", "q_sizes", "k_sizes"))
        assert int(rs.sum(1)[0]) == S and int(cs.sum(1)[0]) == S
        for h in range(H):
            ql = torch.empty(S, dtype=torch.long)
            kl = torch.empty(S, dtype=torch.long)
            ql[qperm[h]] = torch.repeat_interleave(torch.arange(rs.shape[1]), rs[h].long())
            kl[kperm[h]] = torch.repeat_interleave(torch.arange(cs.shape[1]), cs[h].long())
            allowed = m[h][ql][:, kl]
            s = (q[0, h].float() @ k[0, h].float().T) * D ** -0.5
            w = torch.nan_to_num(torch.softmax(s.masked_fill(~allowed, float("-inf")), -1), nan=0.0)
            torch.testing.assert_close(o[0, h], w @ v[0, h].float(), rtol=3e-2, atol=2e-2)


def test_svg1_core_hy_pipeline(cuda):
    """Whole SVG1 core: sample_mse -> argmin -> placement -> band attention -> inverse placement equals the
    oracle pipeline given the same sampled rows."""
    from oracle import attention as oa
    from oracle import layout as ol
    from svgb200.models import hyvideo as hy

    g = torch.Generator().manual_seed(3)
    H, F, P, ctx, plen, D = 4, 5, 200, 56, 30, 128
    S = ctx + F * P
    q, k, v = (torch.randn(1, H, S, D, generator=g).to(torch.bfloat16) for _ in range(3))
    coreobj = hy.HunyuanSVG1Core(ctx, plen, F, P, H, D, 0.45, cuda, num_sampled_rows=32, sample_mse_max_row=500)
    rows = torch.randint(0, 500, (32,), generator=g)
    o = coreobj.sparse_core(q.to(cuda), k.to(cuda), v.to(cuda), sampled_rows=rows).float().cpu()
    masks = [oa.profiling_mask_rows(mn, rows, "hy", ctx, F, P) for mn in ("spatial", "temporal")]
    best = oa.sample_mse(q, k, v, rows, masks).bfloat16().argmin(0).view(-1)
    mul = oa.sparsity_to_width(0.45, ctx, F, P)
    mod = oa.hy_mask_mod(ctx, plen, F, P, mul)
    qp, kp, vp = (ol.head_placement(t[0], best.numpy(), ctx, F, P) for t in (q, k, v))
    ref = ol.head_placement(oa.masked_attention_bhsd(qp, kp, vp, mod).bfloat16(), best.numpy(), ctx, F, P, inverse=True)
    torch.testing.assert_close(o[0], ref.float(), rtol=3e-2, atol=2e-2)
