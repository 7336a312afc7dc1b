"""GPU tests of the BASELINE.json configurations and branches that round 1 left untested:
  * configs[2]  Wan 2.1 720p (S = 75 600, H = 40) through the whole SVG2 core, oracle on sampled rows
  * sample_mse at the HunyuanVideo-720p size against the fp32 oracle
  * the dense branches of attention_core_logic (SVG1 and SVG2 cores), incl. HunyuanVideo's padded-prompt varlen split
    (hyvideo/attention.py:308-316, 452-470, 807-875) and zero_step_kmeans_init
  * sparse_core_from_host called twice with different inputs and no host sync in between
  * CogVideoX SVG1 core (text first) end to end
  * head-parallel output == single-GPU output, bit for bit, on >= 2 GPUs (skipped on a 1-GPU box)
  * two variable-block plans of one shape held at the same time
"""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _oracle_rows(q, k, v, rows, allowed):
    D = q.shape[-1]
    s = (q[rows].float() @ k.float().T) * D ** -0.5
    s = s.masked_fill(~allowed, float("-inf"))
    return torch.nan_to_num(torch.softmax(s, dim=-1), nan=0.0) @ v.float()


def _labels_from(perm, sizes):
    """cluster id of every ORIGINAL token from the engine's sorted indices + cluster sizes"""
    S = perm.shape[0]
    lab = torch.empty(S, dtype=torch.long, device=perm.device)
    lab[perm.long()] = torch.repeat_interleave(torch.arange(sizes.shape[0], device=perm.device), sizes.long())
    return lab


def test_wan720p_svg2_fullsize_pipeline(cuda):
    """BASELINE configs[2]: Wan 2.1 T2V 720p, S = 21 x 3600 = 75 600, 40 heads, QC = 300 / KC = 1000, top-p 0.9
    (scripts/wan/wan_t2v_720p_sap.sh).  The engine's integer outputs define the element mask on the original token
    order; sampled rows of every 8th head are checked against the fp32 oracle, the integers for consistency."""
    from svgb200.models import wan

    H, F, P, D = 40, 21, 3600, 128
    S = F * P
    g = torch.Generator(device=cuda).manual_seed(5)
    q, k, v = (torch.randn(1, H, S, D, device=cuda, generator=g).to(torch.bfloat16) for _ in range(3))
    # give the tokens cluster structure (k-means on pure noise yields near-uniform maps)
    base = torch.randn(1, H, 64, D, device=cuda, generator=g) * 1.5
    which = torch.randint(0, 64, (S,), device=cuda, generator=g)
    q = (q.float() * 0.7 + base[:, :, which]).to(torch.bfloat16)
    k = (k.float() * 0.7 + base[:, :, which]).to(torch.bfloat16)
    sap = wan.WanSAPCore(F, P, num_q_centroids=300, num_k_centroids=1000, top_p_kmeans=0.9, min_kc_ratio=0.1,
                         kmeans_iter_init=4, kmeans_iter_step=2)
    for step in range(2):
        o = sap.sparse_core(q, k, v)
        m, rs, cs = (sap.last[x] for x in ("dynamic_map", "q_sizes", "k_sizes"))
        qperm, kperm = sap.last["q_sorted_indices"], sap.last["k_sorted_indices"]
        assert bool((rs.sum(1) == S).all()) and bool((cs.sum(1) == S).all())
        assert bool((torch.sort(qperm, dim=1).values == torch.arange(S, device=cuda)).all())
        assert bool((torch.sort(kperm, dim=1).values == torch.arange(S, device=cuda)).all())
        gr = torch.Generator().manual_seed(step)
        rows = torch.randint(0, S, (48,), generator=gr).to(cuda)
        for h in range(0, H, 8):
            ql, kl = _labels_from(qperm[h], rs[h]), _labels_from(kperm[h], cs[h])
            allowed = m[h][ql[rows]][:, kl]
            ref = _oracle_rows(q[0, h], k[0, h], v[0, h], rows, allowed)
            torch.testing.assert_close(o[0, h][rows].float(), ref, rtol=3e-2, atol=2e-2)
    dens = (m.float() * rs[:, :, None].float() * cs[:, None, :].float()).sum((1, 2)) / (S * S)
    assert 0.02 < dens.mean().item() < 0.9


def test_sample_mse_hunyuan_fullsize(cuda):
    """sample_mse at S = 119 056 (64 rows < 10 000, hyvideo/inference.py:43-44) vs the fp32 oracle."""
    from oracle import attention as oa
    from svgb200 import core

    ctx, F, P, D, H = 256, 33, 3600, 128, 3
    S = ctx + F * P
    g = torch.Generator(device=cuda).manual_seed(6)
    q, k, v = (torch.randn(H, S, D, device=cuda, generator=g).to(torch.bfloat16) for _ in range(3))
    k[0] = (k[0].float() * 0.2 + q[0].float()).to(torch.bfloat16)  # a head with local structure
    rows = torch.randint(0, 10000, (64,), generator=torch.Generator().manual_seed(7))
    got = core.sample_mse(q, k, v, rows.to(cuda), 0, ctx, F, P).cpu()
    kv = torch.arange(S)
    ref = torch.zeros(2, H)
    for i, mn in enumerate(("spatial", "temporal")):
        allowed = oa.profiling_mask_rows(mn, rows, "hy", ctx, F, P).to(cuda)
        for h in range(H):
            full = _oracle_rows(q[h], k[h], v[h], rows.to(cuda), torch.ones_like(allowed))
            part = _oracle_rows(q[h], k[h], v[h], rows.to(cuda), allowed)
            ref[i, h] = ((part - full) ** 2).mean().item()
    torch.testing.assert_close(got, ref, rtol=3e-2, atol=1e-7)
    assert torch.equal(got.argmin(0), ref.argmin(0))


def _sdpa_segments(q, k, v, segs):
    out = torch.empty_like(q)
    s0 = 0
    for n in segs:
        out[:, :, s0:s0 + n] = torch.nn.functional.scaled_dot_product_attention(
            q[:, :, s0:s0 + n].float(), k[:, :, s0:s0 + n].float(), v[:, :, s0:s0 + n].float()).to(q.dtype)
        s0 += n
    return out


def test_dense_branches_and_padded_prompt_varlen(cuda):
    """dense_attention(seg_lens=[F*P + prompt, pad]) == per-segment attention (the reference's flash_attn_varlen /
    flashinfer_varlen_func with cu_seqlens = [0, sum(mask), S]); both cores take the dense branch when
    layer_idx < first_layers_fp or timestep[0] > first_times_fp (hyvideo/attention.py:489-502, 733-745)."""
    from svgb200.models import hyvideo as hy
    from svgb200.models.common import dense_attention

    H, F, P, ctx, plen, D = 3, 4, 150, 40, 17, 128
    V, S = F * P, ctx + F * P
    g = torch.Generator().manual_seed(8)
    q, k, v = (torch.randn(1, H, S, D, generator=g).to(torch.bfloat16).to(cuda) for _ in range(3))
    segs = [V + plen, ctx - plen]
    ref = _sdpa_segments(q, k, v, segs)
    torch.testing.assert_close(dense_attention(q, k, v, segs).float(), ref.float(), rtol=3e-2, atol=2e-2)
    torch.testing.assert_close(dense_attention(q, k, v).float(), _sdpa_segments(q, k, v, [S]).float(), rtol=3e-2, atol=2e-2)
    cu = torch.tensor([0, V + plen, S], dtype=torch.int32, device=cuda)
    cu_max = (cu, cu, S, S)
    # SVG1 core: timestep above the threshold -> dense; below -> sparse (and the two differ)
    svg1 = hy.HunyuanSVG1Core(ctx, plen, F, P, H, D, 0.3, cuda, num_sampled_rows=16, sample_mse_max_row=400,
                              first_times_fp=900, first_layers_fp=0, layer_idx=2)
    o_dense = svg1.attention_core_logic(q, k, v, torch.tensor([950.0]), 2, cu_max)
    torch.testing.assert_close(o_dense.float(), ref.float(), rtol=3e-2, atol=2e-2)
    o_sparse = svg1.attention_core_logic(q, k, v, torch.tensor([100.0]), 2, cu_max)
    assert (o_sparse.float() - ref.float()).abs().max() > 0.05
    svg1_layer = hy.HunyuanSVG1Core(ctx, plen, F, P, H, D, 0.3, cuda, first_times_fp=900, first_layers_fp=3, layer_idx=2)
    torch.testing.assert_close(svg1_layer.attention_core_logic(q, k, v, torch.tensor([100.0]), 2, cu_max).float(),
                               ref.float(), rtol=3e-2, atol=2e-2)
    # SVG2 core: dense branch + zero_step_kmeans_init warms the centroids up during the dense steps (:739-743)
    sap = hy.HunyuanSAPCore(ctx, F, P, num_q_centroids=6, num_k_centroids=12, top_p_kmeans=0.9, min_kc_ratio=0.1,
                            kmeans_iter_init=3, kmeans_iter_step=1, prompt_length=plen, zero_step_kmeans_init=True,
                            first_times_fp=900)
    assert not sap.state.q_centroids
    torch.testing.assert_close(sap.attention_core_logic(q, k, v, torch.tensor([950.0]), 0, cu_max).float(), ref.float(),
                               rtol=3e-2, atol=2e-2)
    assert 0 in sap.state.q_centroids and sap.state.q_centroids[0].shape == (H, 6, D)
    o2 = sap.attention_core_logic(q, k, v, torch.tensor([100.0]), 0, cu_max)  # sparse, warm-started
    assert torch.isfinite(o2.float()).all()


def test_from_host_pipeline_back_to_back_calls(cuda):
    """sparse_core_from_host twice with DIFFERENT inputs and no host synchronisation in between: the second call's H2D
    copies must not overwrite staging buffers the first call's last head groups still read."""
    from svgb200.models import hyvideo as hy

    H, F, P, ctx, plen, D = 6, 4, 600, 64, 20, 128
    S = ctx + F * P
    core_obj = hy.HunyuanSVG1Core(ctx, plen, F, P, 2, D, 0.4, cuda, num_sampled_rows=16, sample_mse_max_row=1000)
    rows = torch.randint(0, 1000, (16,), generator=torch.Generator().manual_seed(1))
    g = torch.Generator().manual_seed(9)
    ins = [[torch.randn(1, H, S, D, generator=g).to(torch.bfloat16).pin_memory() for _ in range(3)] for _ in range(2)]
    outs = [torch.empty(1, H, S, D, dtype=torch.bfloat16).pin_memory() for _ in range(2)]
    refs = []
    for a in ins:
        refs.append(core_obj.sparse_core(*(t.to(cuda) for t in a), sampled_rows=rows).cpu())
    torch.cuda.synchronize()
    for rep in range(3):
        core_obj.sparse_core_from_host(*ins[0], outs[0], sampled_rows=rows)
        core_obj.sparse_core_from_host(*ins[1], outs[1], sampled_rows=rows)  # no sync in between
        torch.cuda.synchronize()
        assert torch.equal(outs[0], refs[0]) and torch.equal(outs[1], refs[1]), rep


def test_cog_svg1_core_pipeline(cuda):
    """CogVideoX (text FIRST): profiling masks of cog/utils.py:61-88, text-first placement, executed mask of
    cog/utils.py:30-46; equals the oracle pipeline on the same sampled rows."""
    from oracle import attention as oa
    from oracle import layout as ol
    from svgb200.models import cog

    H, F, P, ctx, D = 4, 5, 200, 34, 64
    S = ctx + F * P
    sys.path.insert(0, str(ROOT / "tests" / "golden"))
    from gen_inputs import structured_qkv

    g = torch.Generator().manual_seed(10)
    q, k, v = structured_qkv(10, H, F, P, ctx, D, text_first=True)
    c = cog.CogSVG1Core(ctx, F, P, H, D, 0.45, cuda, num_sampled_rows=32)
    for rows in (torch.randint(ctx, S, (32,), generator=g),            # video rows only
                 torch.cat([torch.tensor([3]), torch.randint(ctx, S, (31,), generator=g)])):  # a text row -> all temporal
        o = c.sparse_core(q.to(cuda), k.to(cuda), v.to(cuda), sampled_rows=rows).float().cpu()
        masks = [oa.profiling_mask_rows_cog(mn, rows, ctx, F, P) for mn in ("spatial", "temporal")]
        mses = oa.sample_mse(q, k, v, rows, masks)
        if bool((rows < ctx).any()):
            assert torch.isnan(mses[1]).all()
        best = torch.argmin(mses.bfloat16(), dim=0).view(-1)
        if bool((rows < ctx).any()):
            assert (best == 1).all()
        mul = oa.sparsity_to_width(0.45, ctx, F, P)
        mod = oa.cog_mask_mod(ctx, F, P, mul)
        qp, kp, vp = (ol.head_placement(t[0], best.numpy(), ctx, F, P, text_first=True) for t in (q, k, v))
        ref = ol.head_placement(oa.masked_attention_bhsd(qp, kp, vp, mod).bfloat16(), best.numpy(), ctx, F, P,
                                text_first=True, inverse=True)
        torch.testing.assert_close(o[0], ref.float(), rtol=3e-2, atol=2e-2)
    # dense switch of the Cog processor (cog/attention.py:172-175)
    c.first_times_fp = 0.2
    dense = c.attention_core_logic(q.to(cuda), k.to(cuda), v.to(cuda), torch.tensor([900.0]))
    torch.testing.assert_close(dense.float().cpu(), _sdpa_segments(q, k, v, [S]).float(), rtol=3e-2, atol=2e-2)


def test_two_plans_of_one_shape_coexist(cuda):
    """Two variable-block plans with identical (BH, S, QC, KC) built first and executed afterwards (the
    shape-keyed workspace cache used to let the second plan overwrite the first one's work list)."""
    from oracle.attention import dynamic_block_sparse_fwd
    from svgb200 import core

    g = torch.Generator().manual_seed(11)
    H, S, D, QC, KC = 2, 1024, 64, 4, 8
    q, k, v = (torch.randn(1, H, S, D, generator=g).to(torch.float16) for _ in range(3))
    row = torch.full((H, QC), S // QC, dtype=torch.int32)
    col = torch.full((H, KC), S // KC, dtype=torch.int32)
    m1 = torch.rand(H, QC, KC, generator=g) < 0.5
    m2 = ~m1
    p1 = core.plan_varblock(m1.to(cuda), row.to(cuda), col.to(cuda), S)
    p2 = core.plan_varblock(m2.to(cuda), row.to(cuda), col.to(cuda), S)
    assert p1.ws.data_ptr() != p2.ws.data_ptr()
    o2 = core.attn_fwd(q.to(cuda), k.to(cuda), v.to(cuda), p2).float().cpu()
    o1 = core.attn_fwd(q.to(cuda), k.to(cuda), v.to(cuda), p1).float().cpu()
    for o, m in ((o1, m1), (o2, m2)):
        ref = dynamic_block_sparse_fwd(q, k, v, m[None], row[None], col[None])
        torch.testing.assert_close(o, ref, rtol=5e-3, atol=5e-3)


@pytest.mark.skipif(torch.cuda.is_available() and torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_head_parallel_equals_single_gpu(cuda):
    """sparse_core_head_parallel over 2 NCCL ranks (multi-stream issue + overlapped all-gather) must reproduce the
    single-GPU sparse_core bit for bit — catches stream-ordering bugs in parallel.run_overlapped."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", str(ROOT / "tests" / "mp_head_parallel.py")],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "HEAD_PARALLEL_OK" in r.stdout
