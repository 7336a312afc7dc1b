"""Parity at BASELINE.json's full sizes (HunyuanVideo 720p: S = 119056, D = 128), where the dense oracle
cannot run: (a) the oracle on a random SAMPLE OF QUERY ROWS against all keys, (b) size-independent
properties — softmax rows sum to one (constant-V test), linearity in V, permutation round trips,
sortedness and bijectivity of the argsort, placement round trips."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CTX, F, P, PLEN, D = 256, 33, 3600, 60, 128
S = CTX + F * P
H = 2  # two heads keep the test in seconds; the kernels treat heads independently


def _qkv(cuda, seed):
    g = torch.Generator(device=cuda).manual_seed(seed)
    return [torch.randn(1, H, S, D, device=cuda, dtype=torch.float32, generator=g).to(torch.bfloat16) for _ in range(3)]


def _oracle_rows(q, k, v, rows, allowed):
    """fp32 masked attention for selected rows.  q [S,D] etc. on the GPU (torch fp32 math as plumbing for the
    oracle formula softmax(q k^T / sqrt(d) + mask) v — same as oracle.attention.ref_torch_attn_impl)."""
    s = (q[rows].float() @ k.float().T) * D ** -0.5
    s = s.masked_fill(~allowed, float("-inf"))
    return torch.softmax(s, dim=-1) @ v.float()


def test_band_fullsize_sampled_rows_and_properties(cuda):
    from oracle import attention as oa
    from svgb200 import core
    from svgb200.models import hyvideo as hy

    q, k, v = _qkv(cuda, 1)
    mul = oa.sparsity_to_width(0.30, CTX, F, P)
    bm = hy.prepare_flexattention(1, H, D, torch.bfloat16, cuda, CTX, PLEN, F, P, diag_width=mul, multiplier=mul)
    o = hy.sparse_flex_attention(q, k, v, bm)
    # (a) oracle on sampled rows, incl. the video/text/padding boundaries and band edges
    g = torch.Generator().manual_seed(0)
    rows = torch.cat([torch.randint(0, S, (48,), generator=g),
                      torch.tensor([0, 127, 128, 19071, 19072, F * P - 1, F * P, F * P + PLEN - 1, F * P + PLEN, S - 1])])
    mod = oa.hy_mask_mod(CTX, PLEN, F, P, mul)
    allowed = mod(rows.view(-1, 1), torch.arange(S).view(1, -1)).to(cuda)
    for h in range(H):
        ref = _oracle_rows(q[0, h], k[0, h], v[0, h], rows.to(cuda), allowed)
        torch.testing.assert_close(o[0, h][rows.to(cuda)].float(), ref, rtol=3e-2, atol=2e-2)
    # (b1) constant V columns -> every output row equals that constant (weights sum to one)
    vc = torch.linspace(-2, 2, D, device=cuda).to(torch.bfloat16).expand(1, H, S, D).contiguous()
    oc = hy.sparse_flex_attention(q, k, vc, bm).float()
    torch.testing.assert_close(oc, vc.float(), rtol=2e-2, atol=2e-2)
    # (b2) linearity in V
    v2 = torch.randn_like(v)
    o2 = hy.sparse_flex_attention(q, k, v2, bm).float()
    o12 = hy.sparse_flex_attention(q, k, (v.float() + v2.float()).to(torch.bfloat16), bm).float()
    torch.testing.assert_close(o12, o.float() + o2, rtol=5e-2, atol=5e-2)


def test_varblock_fullsize_sampled_rows(cuda):
    """QC=400+2 / KC=1000+2 map at 30 % (the SVG2 shape incl. HunyuanVideo's prompt / padding blocks)."""
    from svgb200 import core

    q, k, v = _qkv(cuda, 2)
    QC, KC, V = 400, 1000, F * P
    g = torch.Generator().manual_seed(3)

    def sizes(n):
        cuts = torch.sort(torch.randperm(V - 1, generator=g)[: n - 1] + 1)[0]
        return torch.diff(torch.cat([torch.tensor([0]), cuts, torch.tensor([V])])).to(torch.int32)
    row = torch.stack([torch.cat([sizes(QC), torch.tensor([PLEN, CTX - PLEN], dtype=torch.int32)]) for _ in range(H)])
    col = torch.stack([torch.cat([sizes(KC), torch.tensor([PLEN, CTX - PLEN], dtype=torch.int32)]) for _ in range(H)])
    bm = torch.zeros(H, QC + 2, KC + 2, dtype=torch.bool)
    bm[:, :QC, :KC] = torch.rand(H, QC, KC, generator=g) < 0.3
    bm[:, -2, :-1] = True
    bm[:, :-1, -2] = True
    bm[:, -1, -1] = True
    bm[0, 7, :] = False  # one q-block that sees nothing -> zeros
    plan = core.plan_varblock(bm.to(cuda), row.to(cuda), col.to(cuda), S)
    o = core.attn_fwd(q, k, v, plan)
    rows = torch.randint(0, S, (64,), generator=g)
    for h in range(H):
        rq = torch.repeat_interleave(torch.arange(QC + 2), row[h].long())
        ck = torch.repeat_interleave(torch.arange(KC + 2), col[h].long())
        allowed = bm[h][rq[rows]][:, ck].to(cuda)
        ref = torch.nan_to_num(_oracle_rows(q[0, h], k[0, h], v[0, h], rows.to(cuda), allowed), nan=0.0)
        torch.testing.assert_close(o[0, h][rows.to(cuda)].float(), ref, rtol=1e-2, atol=1e-2)
    r0 = int(row[0, :7].sum())
    assert torch.all(o[0, 0, r0:r0 + int(row[0, 7])] == 0)


def test_layout_roundtrips_fullsize(cuda):
    from svgb200 import core

    Hh, K = 24, 1000
    g = torch.Generator(device=cuda).manual_seed(4)
    x = torch.randn(1, Hh, S, D, device=cuda, generator=g).to(torch.bfloat16)
    labels = torch.randint(0, K, (Hh, S), device=cuda, generator=g)
    perm, counts = core.argsort_labels(labels, K)
    sorted_labels = torch.gather(labels, 1, perm.long())
    assert bool((sorted_labels[:, 1:] >= sorted_labels[:, :-1]).all())                    # sortedness
    assert bool((torch.sort(perm, dim=1).values == torch.arange(S, device=cuda)).all())    # bijection
    same = sorted_labels[:, 1:] == sorted_labels[:, :-1]
    assert bool((perm[:, 1:][same] > perm[:, :-1][same]).all())                            # stability
    ref_counts = torch.zeros(Hh, K, dtype=torch.int64, device=cuda).scatter_add_(1, labels, torch.ones_like(labels))
    assert torch.equal(counts.long(), ref_counts)
    y = core.permute_gather(x, perm)
    assert torch.equal(core.permute_scatter(y, perm), x)                                    # round trip, bit exact
    best = torch.randint(0, 2, (1, Hh), device=cuda, generator=g)
    a, b = torch.empty_like(x), torch.empty_like(x)
    core.head_placement([x], [a], best, CTX, F, P)
    core.head_placement([a], [b], best, CTX, F, P, inverse=True)
    assert torch.equal(b, x)
    assert torch.equal(a[0, best[0] == 0], x[0, best[0] == 0])
