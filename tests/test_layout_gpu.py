"""GPU parity (bit-exact) for the layout transforms: argsort-by-label, permute gather/scatter,
head placement.  Oracles: oracle/layout.py (restating permute.py:12-170, placement.py:156-184,390-401).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("BH,S,K", [(3, 5000, 100), (2, 74256 // 8, 1000), (1, 37, 4), (2, 1024, 1)])
def test_argsort_labels_stable_and_counts(cuda, BH, S, K):
    from oracle.layout import stable_argsort_labels
    from svgb200 import core

    g = torch.Generator().manual_seed(S)
    labels = torch.randint(0, K, (BH, S), generator=g)
    perm, counts = core.argsort_labels(labels.to(cuda), K)
    torch.cuda.synchronize()
    ref = stable_argsort_labels(labels.numpy())
    assert np.array_equal(perm.cpu().numpy(), ref)
    ref_counts = np.stack([np.bincount(labels[b].numpy(), minlength=K) for b in range(BH)])
    assert np.array_equal(counts.cpu().numpy(), ref_counts)


@pytest.mark.parametrize("D", [64, 128])
def test_permute_gather_scatter_roundtrip_and_oracle(cuda, D):
    """permute.py __main__ check (:198-201): permute then inverse == identity; plus oracle equality."""
    from oracle import layout as ol
    from svgb200 import core

    g = torch.Generator().manual_seed(1)
    B, H, S = 1, 5, 9283
    x = torch.randn(B, H, S, D, generator=g).to(torch.float16)
    labels = torch.randint(0, 1000, (B * H, S), generator=g)
    perm, _ = core.argsort_labels(labels.to(cuda), 1000)
    y = core.permute_gather(x.to(cuda), perm)
    x_rec = core.permute_scatter(y, perm)
    torch.cuda.synchronize()
    assert torch.equal(x_rec.cpu(), x)
    ref = ol.permute_gather(x.view(B * H, S, D), perm.cpu().numpy())
    assert torch.equal(y.cpu().view(B * H, S, D), ref)
    assert torch.equal(ol.permute_scatter(ref, perm.cpu().numpy()), x.view(B * H, S, D))


@pytest.mark.parametrize("text_first", [False, True])
def test_head_placement_and_inverse(cuda, text_first):
    """placement.py:187-221 / 404-432 self-checks (bit exact), HY (text last) and Cog (text first)."""
    from oracle import layout as ol
    from svgb200 import core

    g = torch.Generator().manual_seed(2)
    ctx, F, P, cfg, H, D = 226, 11, 408, 2, 6, 64
    S = ctx + F * P
    q, k, v = (torch.randn(cfg, H, S, D, generator=g).to(torch.bfloat16) for _ in range(3))
    best = torch.randint(0, 2, (cfg, H), generator=g)
    outs = [torch.empty_like(q, device=cuda) for _ in range(3)]
    core.head_placement([t.to(cuda) for t in (q, k, v)], outs, best.to(cuda), ctx, F, P, text_first=text_first)
    torch.cuda.synchronize()
    for t, o in zip((q, k, v), outs):
        ref = ol.head_placement(t.view(cfg * H, S, D), best.view(-1).numpy(), ctx, F, P, text_first=text_first)
        assert torch.equal(o.cpu().view(cfg * H, S, D), ref)
    back = [torch.empty_like(q, device=cuda)]
    core.head_placement([outs[0]], back, best.to(cuda), ctx, F, P, text_first=text_first, inverse=True)
    assert torch.equal(back[0].cpu(), q)


def test_density(cuda):
    from oracle.kmeans import density_calculation
    from svgb200 import core

    g = torch.Generator().manual_seed(4)
    H, QC, KC = 5, 40, 100
    m = torch.rand(1, H, QC, KC, generator=g) > 0.6
    r = torch.randint(0, 500, (1, H, QC), generator=g, dtype=torch.int32)
    c = torch.randint(0, 300, (1, H, KC), generator=g, dtype=torch.int32)
    d = core.density(m[0].to(cuda), r[0].to(cuda), c[0].to(cuda))
    torch.testing.assert_close(d.cpu(), density_calculation(m, r, c)[0], rtol=1e-6, atol=1e-7)
