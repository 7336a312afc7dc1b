"""CPU pinning (`-m "not gpu"`) of the SVG1 / ops-API oracles and host mirrors against vectors produced by
EXECUTING the real reference (tests/golden/make_golden_svg1.py -> svg1_golden.npz): HunyuanVideo mask_mod,
get_attention_mask of the three model families, the real processors' sample_mse (bf16), sparsity_to_width,
dynamic_map_post_processing, and the reference's own ops-API test oracles / BSR generators."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE / "golden"))
from gen_inputs import checksum, smse_inputs  # noqa: E402

from oracle import attention as oa  # noqa: E402

G = np.load(HERE / "golden" / "svg1_golden.npz")


def unpack(name, shape):
    n = int(np.prod(shape))
    return torch.from_numpy(np.unpackbits(G[name])[:n].reshape(shape).astype(bool))


@pytest.mark.parametrize("i", [0, 1, 2])
def test_hy_mask_mod_matches_reference(i):
    ctx, plen, F, P, mul = G[f"hy_mm{i}_params"]
    ctx, plen, F, P = int(ctx), int(plen), int(F), int(P)
    S = ctx + F * P
    ref = unpack(f"hy_mm{i}", (S, S))
    qi, ki = torch.arange(S).view(-1, 1), torch.arange(S).view(1, -1)
    assert torch.equal(oa.hy_mask_mod(ctx, plen, F, P, float(mul))(qi, ki), ref)
    # the engine's (mode, m0, m1, m2) parametrisation of the same mask (svgb200.models.hyvideo.band_params)
    assert torch.equal(oa.generic_mask_fn(1, F * P, F * P + plen, oa.hy_band_width(float(mul), P))(qi, ki), ref)


def test_sparsity_to_width_matches_reference():
    got = [oa.sparsity_to_width(0.25, 256, 33, 3600), oa.sparsity_to_width(0.30, 256, 33, 3600),
           oa.sparsity_to_width(0.4, 64, 5, 1000)]
    np.testing.assert_allclose(got, G["hy_s2w"], rtol=1e-13)


@pytest.mark.parametrize("name", ["hy", "hy2", "wan", "wan2", "cog", "cog2"])
@pytest.mark.parametrize("mask_name", ["spatial", "temporal"])
def test_profiling_masks_match_get_attention_mask(name, mask_name):
    ctx, F, P, max_row, nrows, S = (int(x) for x in G[f"prof_{name}_dims"])
    ref = unpack(f"prof_{name}_{mask_name}", (nrows, S))
    rows = list(range(nrows))
    if name.startswith("cog"):
        got = oa.profiling_mask_rows_cog(mask_name, rows, ctx, F, P)
    else:
        got = oa.profiling_mask_rows(mask_name, rows, "hy" if name.startswith("hy") else "wan", ctx, F, P)
    assert torch.equal(got, ref)


def _engine_mask(mode, q, kv, m0, m1, m2):
    """attn_common.cuh:mask_allowed restated 1:1 for the profiling modes (checked against the same goldens, so the
    CUDA predicate, the oracle and the reference are the same function)."""
    d = (q - kv).abs()
    if mode in (8, 9):
        thres, ctx = m2 & 0xfff, m2 >> 12
        if mode == 8:
            lim = ((m0 * m1 + 127) // 128) * 128
            return (q < ctx) | (kv < ctx) | ((q < lim) & (kv < lim) & ((q // 128 - kv // 128).abs() < thres))
        F, P = m0, m1
        qv, kvv = (q - ctx).clamp(min=0), (kv - ctx).clamp(min=0)
        bd = (((qv % P) * F + qv // P) // 128 - ((kvv % P) * F + kvv // P) // 128).abs()
        return (q >= ctx) & (kv >= ctx) & (bd < thres)
    F, P, V = m0, m1, m0 * m1
    hy = mode <= 5
    temporal = mode in (5, 7)
    qc, kc = q.clamp(max=V - 1), kv.clamp(max=V - 1)
    qi = (qc % P) * F + qc // P if temporal else qc
    ki = (kc % P) * F + kc // P if temporal else kc
    inner = ((qi // 128 - ki // 128).abs() < m2) | ((not hy) & (ki < P))
    text = (q >= V) | (kv >= V)
    return torch.where(text, torch.full_like(inner, hy), inner)


@pytest.mark.parametrize("name,modes", [("hy2", (4, 5)), ("wan2", (6, 7)), ("cog2", (8, 9))])
def test_engine_profiling_predicate_matches_reference(name, modes):
    ctx, F, P, max_row, nrows, S = (int(x) for x in G[f"prof_{name}_dims"])
    thres = int((P * (2 if name.startswith("wan") else 1.5)) // 128)
    m2 = thres | (ctx << 12) if name.startswith("cog") else thres
    q, kv = torch.arange(nrows).view(-1, 1), torch.arange(S).view(1, -1)
    for mode, mn in zip(modes, ("spatial", "temporal")):
        assert torch.equal(_engine_mask(mode, q, kv, F, P, m2), unpack(f"prof_{name}_{mn}", (nrows, S))), (name, mn)


@pytest.mark.parametrize("case", ["hy0", "hy1", "wan0", "wan1", "cog0", "cog1", "cog2", "cog3"])
def test_oracle_sample_mse_matches_reference_processor(case):
    """fp32 oracle vs the reference processor's bf16 sample_mse: MSE within bf16 noise of the reference's own
    arithmetic, and the SAME best_mask_idx wherever the two MSEs differ by more than that noise."""
    seed, cfg, H, S, D, csum = G[f"smse_{case}_in"]
    q, k, v = smse_inputs(seed, int(cfg), int(H), int(S), int(D))
    assert abs(checksum(q, k, v) - csum) <= 1e-6 * abs(csum), "seeded inputs differ from the generator's"
    rows = torch.from_numpy(G[f"smse_{case}_rows"])
    ref = torch.from_numpy(G[f"smse_{case}_mses"])
    fam = case[:-1]
    if fam == "cog":
        masks = [oa.profiling_mask_rows_cog(mn, rows, 30, 3, 200) for mn in ("spatial", "temporal")]
    elif fam == "hy":
        masks = [oa.profiling_mask_rows(mn, rows, "hy", 16, 3, 200) for mn in ("spatial", "temporal")]
    else:
        masks = [oa.profiling_mask_rows(mn, rows, "wan", 0, 4, 150) for mn in ("spatial", "temporal")]
    mine = oa.sample_mse(q, k, v, rows, masks)
    nan_ref = torch.isnan(ref)
    if fam == "cog" and bool((rows < 30).any()):
        # the reference's temporal mask leaves text rows empty -> NaN MSE for every head (cog/utils.py:76-86)
        assert nan_ref[1].all() and not nan_ref[0].any()
        assert torch.isnan(mine[1]).all()
        assert (torch.from_numpy(G[f"smse_{case}_best"]) == 1).all()
    else:
        assert not nan_ref.any()
    ok = ~nan_ref
    torch.testing.assert_close(mine[ok], ref[ok], rtol=6e-2, atol=1e-5)
    best_ref = torch.from_numpy(G[f"smse_{case}_best"])
    decided = ((ref[0] - ref[1]).abs() > 0.1 * torch.maximum(ref[0], ref[1])) | nan_ref.any(0)
    mine_best = torch.argmin(mine, dim=0)
    assert torch.equal(mine_best[decided], best_ref[decided])


def test_dynamic_map_post_processing_matches_reference():
    """The HunyuanVideo prompt / padding blocks (hyvideo/attention.py:657-702) as built by SAPCore.sparse_core."""
    H, V, ctx, plen, QC, KC, D = (int(x) for x in G["pp_dims"])
    dyn = torch.from_numpy(G["pp_dyn"])[0]
    # same construction as svgb200/models/common.py:SAPCore.sparse_core (host logic, restated on CPU tensors)
    d2 = torch.nn.functional.pad(dyn, (0, 2, 0, 2), value=False)
    d2[:, -2, :-1] = True
    d2[:, :-1, -2] = True
    d2[:, -1, -1] = True
    assert torch.equal(d2, torch.from_numpy(G["pp_out_dyn"])[0].bool())
    extra = torch.tensor([plen, ctx - plen], dtype=torch.int32).expand(H, 2)
    assert torch.equal(torch.cat([torch.from_numpy(G["pp_qsz"])[0], extra], 1), torch.from_numpy(G["pp_out_qsz"])[0])
    assert torch.equal(torch.cat([torch.from_numpy(G["pp_ksz"])[0], extra], 1), torch.from_numpy(G["pp_out_ksz"])[0])
    tail = torch.arange(V, V + ctx, dtype=torch.int32).expand(H, ctx)
    assert torch.equal(torch.cat([torch.from_numpy(G["pp_qidx"]), tail], 1), torch.from_numpy(G["pp_out_qidx"])[0])
    # the reference writes the permuted video back in front of the text; our engine gathers Q through the index
    # vector instead: gathering the ORIGINAL q by the extended index must give the reference's rewritten tensor when
    # q_perm is the gather of q's video part
    q = torch.from_numpy(G["pp_q"])
    idx = torch.from_numpy(G["pp_out_qidx"])[0].long()
    q_ref = q.clone()
    for h in range(H):
        q_ref[0, h, :V] = q[0, h, idx[h, :V]]
        assert torch.equal(q[0, h, idx[h]], q_ref[0, h])


def test_ops_oracles_match_reference_tests():
    F, P, mul = G["ops_params"]
    F, P = int(F), int(P)
    bm, bs = oa.ref_gen_temporal_mask(F, P, float(mul))
    assert np.array_equal(bm, G["ops_ref_temporal"]) and bs == (P // 10, P // 10)
    bm2, bs2 = oa.ref_gen_spatial_mask(F, P, 1)
    assert np.array_equal(bm2, G["ops_ref_spatial"]) and bs2 == (P, P)
    shape = tuple(int(x) for x in G["ops_b2e_shape"])
    em = oa.gen_mask_block2element(bm, bs, 7)
    assert torch.equal(em, unpack("ops_b2e", shape))
    o = oa.ref_torch_attn_impl(torch.from_numpy(G["ops_attn_q"]), torch.from_numpy(G["ops_attn_k"]),
                               torch.from_numpy(G["ops_attn_v"]), em)
    torch.testing.assert_close(o, torch.from_numpy(G["ops_attn_o"]), rtol=1e-5, atol=1e-6)
    Fw, Pw, mw = G["opsw_params"]
    bmw, bsw = oa.ref_gen_temporal_mask_wan(int(Fw), int(Pw), float(mw), first_frame=False)
    assert np.array_equal(bmw, G["opsw_ref"]) and bsw == (240, 240)


def _bsr_eq(mine, prefix):
    ip, ix, bs = mine
    assert np.array_equal(ip.cpu().numpy(), G[prefix + "_indptr"])
    assert np.array_equal(ix.cpu().numpy(), G[prefix + "_indices"])  # incl. the 256 padding zeros
    assert tuple(bs) == tuple(int(x) for x in G[prefix + "_shape"])


def test_ops_bsr_generators_match_reference():
    """The mirrored BSR generators are pure host code: exact equality with the reference's outputs."""
    sys.path.insert(0, str(HERE.parent / "sparse-videogen_b200"))
    from svgb200.models import wan as wan_m
    from svgb200.ops import _gen_spatial_mask, _gen_temporal_mask, gen_temporal_mask

    F, P, mul = G["ops_params"]
    _bsr_eq(_gen_temporal_mask(int(F), int(P), float(mul), device="cpu"), "ops_bsr_t")
    _bsr_eq(_gen_spatial_mask(int(F), int(P), 1, device="cpu"), "ops_bsr_s")
    Fw, Pw, mw = G["opsw_params"]
    _bsr_eq(gen_temporal_mask(int(Fw), int(Pw), float(mw), device="cpu"), "opsw_bsr")      # ops: band only
    _bsr_eq(wan_m.gen_temporal_mask(int(Fw), int(Pw), float(mw), device="cpu"), "wanu_bsr")  # wan/utils: + first frame
    assert not np.array_equal(G["opsw_bsr_indices"], G["wanu_bsr_indices"])
