"""Host logic of the SVG2 and SVG1 dispatchers on the CPU (no GPU, no CUDA library).

`SAPCore.sparse_core` (svgb200/models/common.py; reference: svg/models/hyvideo/attention.py:555-804, wan :375-559) is
host code around six device operators.  Here those operators are replaced by oracle-backed stand-ins -- test
infrastructure, wired in with monkeypatch -- so the glue itself runs on CPU tensors: the strided video-part views
handed to k-means, the member order returned by the Lloyd loop, the prompt / padding post-processing of map, sizes and
permutations (hyvideo/attention.py:657-702), the centroid warm start across calls and the fused inverse permutation.
The result is compared with plain masked attention under the element mask implied by (labels, map, text rule).
"""
import numpy as np
import pytest
import torch

from oracle import kmeans as ok
from oracle import layout as ol


class _FakePlan:
    def __init__(self, dyn, row_sz, col_sz, S):
        self.dyn, self.row_sz, self.col_sz, self.S = dyn, row_sz, col_sz, S


@pytest.fixture
def cpu_ops(monkeypatch):
    """svgb200.core entry points used by SAPCore.sparse_core -> CPU stand-ins built from the oracle."""
    from svgb200 import core

    calls = {"kmeans_run": [], "argsort_labels": 0}

    def kmeans_run(x, init_centroids, max_iters, tol=1e-4, want_perm=False):
        calls["kmeans_run"].append({"stride0": x.stride(0), "shape": tuple(x.shape), "contiguous": x.is_contiguous(),
                                    "iters": int(max_iters)})
        labels, cent, sizes, n_it = ok.batch_kmeans_euclid(x, init_centroids.shape[1], max_iters, tol,
                                                           init_centroids=init_centroids)
        out = (labels.to(torch.int32), cent, sizes, torch.tensor([n_it], dtype=torch.int32))
        if want_perm:
            perm = torch.from_numpy(np.stack([ol.stable_argsort_labels(l.numpy()) for l in labels])).to(torch.int32)
            return (*out, perm)
        return out

    def dynamic_map(qc, kc, k_sizes, top_p, preserve):
        KC = kc.shape[1]
        ratio = (preserve + 0.5) / KC if preserve > 0 else 0.0  # int(ratio * KC) == preserve
        q_sizes = torch.ones(qc.shape[0], qc.shape[1], dtype=torch.int32)
        return ok.identify_dynamic_map(qc[None], kc[None], q_sizes[None], k_sizes[None], top_p, ratio)[0]

    def permute_gather(t, perm):
        idx = perm.long()[None, :, :, None].expand(1, -1, -1, t.shape[-1])
        return torch.gather(t, 2, idx)

    def plan_varblock(dyn, row_sz, col_sz, S, **_):
        return _FakePlan(dyn, row_sz, col_sz, S)

    def attn_fwd(qp, kp, vp, plan, o_rows=None, **_):
        H, S, D = qp.shape[1], qp.shape[2], qp.shape[3]
        out = torch.zeros_like(qp)
        for h in range(H):
            rl = torch.repeat_interleave(torch.arange(plan.row_sz.shape[1]), plan.row_sz[h].long())
            cl = torch.repeat_interleave(torch.arange(plan.col_sz.shape[1]), plan.col_sz[h].long())
            allowed = plan.dyn[h][rl][:, cl]
            s = (qp[0, h].float() @ kp[0, h].float().T) * D ** -0.5
            w = torch.nan_to_num(torch.softmax(s.masked_fill(~allowed, float("-inf")), -1), nan=0.0)
            o = (w @ vp[0, h].float()).to(qp.dtype)
            if o_rows is None:
                out[0, h] = o
            else:
                out[0, h, o_rows[h].long()] = o  # the fused inverse permutation: row i of the result goes to o_rows[i]
        return out

    def argsort_labels(*a, **k):
        calls["argsort_labels"] += 1
        raise AssertionError("sparse_core must take the member order from the Lloyd loop, not sort again")

    for name, fn in (("kmeans_run", kmeans_run), ("dynamic_map", dynamic_map), ("permute_gather", permute_gather),
                     ("plan_varblock", plan_varblock), ("attn_fwd", attn_fwd), ("argsort_labels", argsort_labels)):
        monkeypatch.setattr(core, name, fn)
    return calls


def _reference(q, k, v, sap, ctx, plen):
    """Masked attention in the ORIGINAL token order under the mask the reference defines: video x video through the
    cluster map; prompt rows/columns see everything but the padding; padding sees only itself."""
    H, S, D = q.shape[1], q.shape[2], q.shape[3]
    V = S - ctx
    last = sap.last
    out = torch.empty(1, H, S, D)
    for h in range(H):
        rs, cs = last["q_sizes"][h].long(), last["k_sizes"][h].long()
        ql = torch.empty(S, dtype=torch.long)
        kl = torch.empty(S, dtype=torch.long)
        ql[last["q_sorted_indices"][h].long()] = torch.repeat_interleave(torch.arange(rs.numel()), rs)
        kl[last["k_sorted_indices"][h].long()] = torch.repeat_interleave(torch.arange(cs.numel()), cs)
        allowed = last["dynamic_map"][h][ql][:, kl]
        if ctx:
            QC, KC = rs.numel() - 2, cs.numel() - 2
            # the text blocks are appended in token order: prompt = block QC / KC, padding = block QC+1 / KC+1
            assert torch.all(ql[V:V + plen] == QC) and torch.all(ql[V + plen:] == QC + 1)
            assert torch.all(kl[V:V + plen] == KC) and torch.all(kl[V + plen:] == KC + 1)
            want = allowed.clone()
            want[V:V + plen, :V + plen] = True
            want[V:V + plen, V + plen:] = False
            want[:V + plen, V:V + plen] = True
            want[:V, V + plen:] = False
            want[V + plen:, :V + plen] = False
            want[V + plen:, V + plen:] = True
            assert torch.equal(allowed, want), "prompt / padding post-processing (hyvideo/attention.py:657-702)"
        s = (q[0, h].float() @ k[0, h].float().T) * D ** -0.5
        w = torch.nan_to_num(torch.softmax(s.masked_fill(~allowed, float("-inf")), -1), nan=0.0)
        out[0, h] = w @ v[0, h].float()
    return out


def test_hunyuan_sap_core_host_logic(cpu_ops):
    from svgb200.models import hyvideo as hy

    g = torch.Generator().manual_seed(7)
    H, F, P, ctx, plen, D = 2, 3, 40, 24, 10, 64
    V, S = F * P, ctx + F * P
    q, k, v = (torch.randn(1, H, S, D, generator=g).to(torch.bfloat16) for _ in range(3))
    sap = hy.HunyuanSAPCore(ctx, F, P, num_q_centroids=4, num_k_centroids=6, top_p_kmeans=0.7, min_kc_ratio=0.2,
                            kmeans_iter_init=3, kmeans_iter_step=1, prompt_length=plen)
    torch.manual_seed(0)  # the first call draws its initial centroids with torch.randint
    for step in range(2):
        o = sap.sparse_core(q, k, v)
        runs = cpu_ops["kmeans_run"][-2:]
        # the video part is clustered in place: a view whose heads are S*D apart, never a packed copy
        assert all(r["shape"] == (H, V, D) and r["stride0"] == S * D and not r["contiguous"] for r in runs)
        assert all(r["iters"] == (3 if step == 0 else 1) for r in runs)  # init, then warm start from the stored centroids
        last = sap.last
        assert last["dynamic_map"].shape == (H, 4 + 2, 6 + 2)
        assert last["q_sorted_indices"].shape == (H, S) and last["q_sorted_indices"].dtype == torch.int32
        assert torch.equal(last["q_sizes"][:, -2:], torch.tensor([[plen, ctx - plen]] * H, dtype=torch.int32))
        assert int(last["q_sizes"].sum(1)[0]) == S and int(last["k_sizes"].sum(1)[0]) == S
        for h in range(H):  # both orders are permutations of all tokens, text tokens in place at the end
            assert torch.equal(torch.sort(last["q_sorted_indices"][h].long()).values, torch.arange(S))
            assert torch.equal(last["k_sorted_indices"][h, V:].long(), torch.arange(V, S))
        ref = _reference(q, k, v, sap, ctx, plen)
        torch.testing.assert_close(o.float(), ref, rtol=3e-2, atol=2e-2)
    assert cpu_ops["argsort_labels"] == 0
    assert len(sap.state.q_centroids) == 1  # one layer key, overwritten by the second call


def test_wan_sap_core_host_logic(cpu_ops):
    from svgb200.models import wan

    g = torch.Generator().manual_seed(8)
    H, F, P, D = 3, 2, 48, 64
    S = F * P
    q, k, v = (torch.randn(1, H, S, D, generator=g).to(torch.bfloat16) for _ in range(3))
    sap = wan.WanSAPCore(F, P, num_q_centroids=3, num_k_centroids=5, top_p_kmeans=0.8, min_kc_ratio=0.0,
                         kmeans_iter_init=2, kmeans_iter_step=2)
    torch.manual_seed(1)
    o = sap.sparse_core(q, k, v)
    runs = cpu_ops["kmeans_run"]
    assert all(r["shape"] == (H, S, D) for r in runs)  # no text: the whole sequence, as it is
    assert sap.last["dynamic_map"].shape == (H, 3, 5)
    torch.testing.assert_close(o.float(), _reference(q, k, v, sap, 0, 0), rtol=3e-2, atol=2e-2)


# ------------------------------------------------------------------------------------------------ SVG1 dispatcher
@pytest.fixture
def cpu_ops_svg1(monkeypatch):
    """Device operators of the SVG1 core -> oracle stand-ins (band plan, profiling, placement, attention)."""
    from oracle import attention as oa
    from svgb200 import core

    class _Band:
        def __init__(self, mode, m0, m1, m2, S):
            self.mode, self.m0, self.m1, self.m2, self.S = mode, m0, m1, m2, S

    calls = {"attn": [], "placement": []}

    def plan_band(mode, m0, m1, m2, BH, S, device):
        return _Band(mode, m0, m1, m2, S)

    def sample_mse(q, k, v, rows, layout, ctx, F, P):
        if layout == 2:  # CogVideoX; rows without keys give 0 in the kernel, NaN in the naive formulation
            masks = [oa.profiling_mask_rows_cog(m, rows, ctx, F, P) for m in ("spatial", "temporal")]
            return torch.nan_to_num(oa.sample_mse(q[None], k[None], v[None], rows, masks)[:, 0], nan=0.0)
        name = {0: "hy", 1: "wan"}[layout]
        masks = [oa.profiling_mask_rows(m, rows, name, ctx, F, P) for m in ("spatial", "temporal")]
        return oa.sample_mse(q[None], k[None], v[None], rows, masks)[:, 0]

    def head_placement(ins, outs, best_mask_idx, ctx, F, P, *, text_first=False, inverse=False):
        calls["placement"].append({"n": len(ins), "inverse": inverse, "text_first": text_first})
        idx = best_mask_idx.reshape(-1).numpy()
        for src, dst in zip(ins, outs):
            dst.copy_(ol.head_placement(src[0], idx, ctx, F, P, text_first=text_first, inverse=inverse)[None])
        return outs

    def plan_varblock(bm, row, col, S, **_):  # dense fall-back = one all-true block
        return ("dense", bm, row, col)

    def attn_fwd(q, k, v, plan, **_):
        if isinstance(plan, tuple):
            calls["attn"].append("dense")
            return oa.masked_attention_bhsd(q[0], k[0], v[0], None)[None].to(q.dtype)
        calls["attn"].append("band")
        fn = oa.generic_mask_fn(plan.mode, plan.m0, plan.m1, plan.m2)
        return oa.masked_attention_bhsd(q[0], k[0], v[0], fn)[None].to(q.dtype)

    for name, fn in (("plan_band", plan_band), ("sample_mse", sample_mse), ("head_placement", head_placement),
                     ("plan_varblock", plan_varblock), ("attn_fwd", attn_fwd)):
        monkeypatch.setattr(core, name, fn)
    return calls


def test_hunyuan_svg1_core_host_logic(cpu_ops_svg1):
    from oracle import attention as oa
    from svgb200.models import hyvideo as hy

    g = torch.Generator().manual_seed(9)
    H, F, P, ctx, plen, D = 3, 4, 128, 40, 17, 64
    S = ctx + F * P
    q, k, v = (torch.randn(1, H, S, D, generator=g).to(torch.bfloat16) for _ in range(3))
    coreobj = hy.HunyuanSVG1Core(ctx, plen, F, P, H, D, 0.4, torch.device("cpu"), num_sampled_rows=16,
                                 sample_mse_max_row=300, first_layers_fp=1, first_times_fp=900, layer_idx=2)
    rows = torch.randint(0, 300, (16,), generator=g)
    o = coreobj.sparse_core(q, k, v, sampled_rows=rows)
    assert cpu_ops_svg1["attn"] == ["band"]
    # one placement call for Q, K, V together, one inverse call for the output
    assert [(c["n"], c["inverse"]) for c in cpu_ops_svg1["placement"]] == [(3, False), (1, True)]
    masks = [oa.profiling_mask_rows(mn, rows, "hy", ctx, F, P) for mn in ("spatial", "temporal")]
    best = oa.sample_mse(q, k, v, rows, masks).bfloat16().argmin(0).view(-1)
    mul = oa.sparsity_to_width(0.4, ctx, F, P)
    mod = oa.hy_mask_mod(ctx, plen, F, P, mul)
    qp, kp, vp = (ol.head_placement(t[0], best.numpy(), ctx, F, P) for t in (q, k, v))
    ref = ol.head_placement(oa.masked_attention_bhsd(qp, kp, vp, mod).bfloat16(), best.numpy(), ctx, F, P, inverse=True)
    torch.testing.assert_close(o[0].float(), ref.float(), rtol=3e-2, atol=2e-2)

    # dispatcher (hyvideo/attention.py:473-524): dense for the first layers and for the early (large) timesteps
    cpu_ops_svg1["attn"].clear()
    coreobj.attention_core_logic(q, k, v, timestep=torch.tensor([950]))
    assert cpu_ops_svg1["attn"] == ["dense"]
    cpu_ops_svg1["attn"].clear()
    coreobj.attention_core_logic(q, k, v, timestep=torch.tensor([500]))
    assert cpu_ops_svg1["attn"] == ["band"]
    coreobj.layer_idx = 0
    cpu_ops_svg1["attn"].clear()
    coreobj.attention_core_logic(q, k, v, timestep=torch.tensor([500]))
    assert cpu_ops_svg1["attn"] == ["dense"]


def test_cog_svg1_core_host_logic(cpu_ops_svg1):
    """CogVideoX: text FIRST, rows sampled over the whole sequence, a sampled text row makes every head temporal (the
    reference's NaN argmin, cog/utils.py:76-86 + cog/attention.py:124-160), scaled dense thresholds (:172-175)."""
    from oracle import attention as oa
    from svgb200.models import cog

    g = torch.Generator().manual_seed(12)
    H, F, P, ctx, D = 2, 3, 128, 26, 64
    S = ctx + F * P
    q, k, v = (torch.randn(1, H, S, D, generator=g).to(torch.bfloat16) for _ in range(3))
    c = cog.CogSVG1Core(ctx, F, P, H, D, 0.4, torch.device("cpu"), num_sampled_rows=12, first_layers_fp=0.05,
                        first_times_fp=0.2, layer_idx=10)
    for rows in (torch.randint(ctx, S, (12,), generator=g),
                 torch.cat([torch.tensor([5]), torch.randint(ctx, S, (11,), generator=g)])):
        cpu_ops_svg1["placement"].clear()
        o = c.sparse_core(q, k, v, sampled_rows=rows)
        assert all(call["text_first"] for call in cpu_ops_svg1["placement"])
        masks = [oa.profiling_mask_rows_cog(mn, rows, ctx, F, P) for mn in ("spatial", "temporal")]
        best = torch.argmin(oa.sample_mse(q, k, v, rows, masks).bfloat16(), dim=0).view(-1)
        if bool((rows < ctx).any()):
            assert (best == 1).all()
        mod = oa.cog_mask_mod(ctx, F, P, oa.sparsity_to_width(0.4, ctx, F, P))
        qp, kp, vp = (ol.head_placement(t[0], best.numpy(), ctx, F, P, text_first=True) for t in (q, k, v))
        ref = ol.head_placement(oa.masked_attention_bhsd(qp, kp, vp, mod).bfloat16(), best.numpy(), ctx, F, P,
                                text_first=True, inverse=True)
        torch.testing.assert_close(o[0].float(), ref.float(), rtol=3e-2, atol=2e-2)
    # dense while layer_idx < 42 * first_layers_fp (= 2.1) or timestep > 1000 * (1 - first_times_fp) (= 800)
    for layer, ts, want in ((10, 700.0, "band"), (10, 900.0, "dense"), (2, 700.0, "dense"), (3, 700.0, "band")):
        c.layer_idx = layer
        cpu_ops_svg1["attn"].clear()
        c.attention_core_logic(q, k, v, torch.tensor([ts]))
        assert cpu_ops_svg1["attn"] == [want], (layer, ts)


def test_wan_svg1_core_host_logic(cpu_ops_svg1):
    """Wan 2.1 (no text; first-frame sink + band, wan/utils.py:30-46; profiling layout 1)."""
    from oracle import attention as oa
    from svgb200.models import wan

    g = torch.Generator().manual_seed(13)
    H, F, P, D = 2, 4, 128, 64
    S = F * P
    q, k, v = (torch.randn(1, H, S, D, generator=g).to(torch.bfloat16) for _ in range(3))
    c = wan.WanSVG1Core(F, P, H, D, 0.5, torch.device("cpu"), num_sampled_rows=16, sample_mse_max_row=400)
    rows = torch.randint(0, 400, (16,), generator=g)
    o = c.sparse_core(q, k, v, sampled_rows=rows)
    masks = [oa.profiling_mask_rows(mn, rows, "wan", 0, F, P) for mn in ("spatial", "temporal")]
    best = oa.sample_mse(q, k, v, rows, masks).bfloat16().argmin(0).view(-1)
    mod = oa.wan_mask_mod(F, P, oa.sparsity_to_width(0.5, 0, F, P))
    qp, kp, vp = (ol.head_placement(t[0], best.numpy(), 0, F, P) for t in (q, k, v))
    ref = ol.head_placement(oa.masked_attention_bhsd(qp, kp, vp, mod).bfloat16(), best.numpy(), 0, F, P, inverse=True)
    torch.testing.assert_close(o[0].float(), ref.float(), rtol=3e-2, atol=2e-2)
