"""Generate golden vectors from the REAL reference (runs only in the build container, where
/root/reference is mounted).  The committed .npz files pin oracle/ on the GPU box, where the reference
does not exist.

    python tests/golden/make_golden.py

Covers every reference function on the hot path that runs on CPU:
  svg.kmeans_utils.dynamic_block_sparse_fwd_torch   (kmeans_utils.py:902-995)
  svg.kmeans_utils.identify_dynamic_map             (:864-896)   [unstable sort: ties may differ]
  svg.kmeans_utils.density_calculation              (:13-31)
  svg.kmeans_utils.permute_tensor_by_labels / apply_inverse_permutation (:820-849)
  svg.models.hyvideo.placement.ref_* / svg.models.cog.placement.ref_*   (placement.py:156-184,390-401)
  svg.models.wan.utils.generate_temporal_head_mask_mod, sparsity_to_width (wan/utils.py:25-61)
  svg.models.cog.utils.generate_temporal_head_mask_mod                   (cog/utils.py:30-46)
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from ref_import import import_kmeans_utils, import_placement  # noqa: E402


def main():
    ku = import_kmeans_utils()
    g = torch.Generator().manual_seed(0)
    out = {}

    # ---- variable-block attention (fp32 so the golden is dtype-independent)
    B, H, S, D, QC, KC = 1, 2, 96, 16, 4, 7
    q, k, v = (torch.randn(B, H, S, D, generator=g) for _ in range(3))
    def part(n):
        cuts = torch.sort(torch.randperm(S - 1, generator=g)[: n - 1] + 1)[0]
        return torch.diff(torch.cat([torch.tensor([0]), cuts, torch.tensor([S])]))
    qs = torch.stack([part(QC) for _ in range(H)])[None]
    ks = torch.stack([part(KC) for _ in range(H)])[None]
    m = torch.rand(B, H, QC, KC, generator=g) > 0.5
    m[0, 0, 1, :] = False
    o = ku.dynamic_block_sparse_fwd_torch(q, k, v, m, qs, ks)
    out.update(vb_q=q.numpy(), vb_k=k.numpy(), vb_v=v.numpy(), vb_map=m.numpy(), vb_qs=qs.numpy(), vb_ks=ks.numpy(),
               vb_o=o.numpy())

    # ---- dynamic map (bf16 inputs as in the live path) + density
    qc = torch.randn(1, 3, 12, 32, generator=g).bfloat16()
    kc = torch.randn(1, 3, 40, 32, generator=g).bfloat16()
    kcs = torch.randint(0, 50, (1, 3, 40), generator=g, dtype=torch.int32)
    qcs = torch.randint(1, 50, (1, 3, 12), generator=g, dtype=torch.int32)
    dm = ku.identify_dynamic_map(qc, kc, qcs, kcs, 0.9, 0.1)
    probs = ku.weighted_softmax(torch.matmul(qc, kc.transpose(-2, -1)) / (32 ** 0.5), kcs.unsqueeze(-2).float())
    dens = ku.density_calculation(dm, qcs, kcs)
    out.update(dm_qc=qc.float().numpy(), dm_kc=kc.float().numpy(), dm_ks=kcs.numpy(), dm_qs=qcs.numpy(),
               dm_map=dm.numpy(), dm_probs=probs.float().numpy(), dm_density=dens.numpy())

    # ---- permutation (torch reference versions)
    x = torch.randn(1, 2, 50, 8, generator=g)
    labels = torch.randint(0, 5, (1, 2, 50), generator=g)
    xp, idx = ku.permute_tensor_by_labels(x, labels, dim=2)
    xr = ku.apply_inverse_permutation(xp, idx, dim=2)
    out.update(pm_x=x.numpy(), pm_labels=labels.numpy(), pm_xp=xp.numpy(), pm_idx=idx.numpy(), pm_xr=xr.numpy())

    # ---- placement (HY text-last, Cog text-first)
    for name, model in (("hy", "hyvideo"), ("cog", "cog")):
        pl = import_placement(model)
        ctx, F, P, cfg, Hh, Dd = 6, 3, 10, 2, 4, 8
        S2 = ctx + F * P
        qq, kk, vv = (torch.randn(cfg, Hh, S2, Dd, generator=g) for _ in range(3))
        best = torch.randint(0, 2, (cfg, Hh), generator=g)
        fwd = getattr(pl, "ref_hunyuan_sparse_head_placement", None) or getattr(pl, "ref_sparse_head_placement")
        inv = getattr(pl, "ref_hunyuan_hidden_states_placement", None) or getattr(pl, "ref_hidden_states_placement")
        qo, ko, vo = fwd(qq.clone(), kk.clone(), vv.clone(), best, ctx, F, P)
        back = torch.zeros_like(qq)
        inv(qo.clone(), back, best, ctx, F, P)
        out.update({f"pl_{name}_q": qq.numpy(), f"pl_{name}_best": best.numpy(), f"pl_{name}_qo": qo.numpy(),
                    f"pl_{name}_back": back.numpy(), f"pl_{name}_dims": np.array([ctx, F, P])})

    # ---- mask mods + sparsity_to_width
    sys.path.insert(0, "/root/reference")
    import importlib
    wu = importlib.import_module("svg.models.wan.utils")
    cu = importlib.import_module("svg.models.cog.utils")
    qi = torch.arange(0, 700).view(-1, 1)
    ki = torch.arange(0, 700).view(1, -1)
    out["mm_wan"] = wu.generate_temporal_head_mask_mod(0, 0, 5, 140, mul=1.3)(None, None, qi, ki).numpy()
    out["mm_cog"] = cu.generate_temporal_head_mask_mod(30, 5, 134, mul=1.2)(None, None, qi, ki).numpy()
    out["mm_cog_sink"] = cu.generate_temporal_head_mask_mod(30, 5, 134, mul=1.2, attn_sink=True)(None, None, qi, ki).numpy()
    out["s2w"] = np.array([wu.sparsity_to_width(0.3, 0, 21, 3600), wu.sparsity_to_width(0.25, 256, 33, 3600),
                           cu.sparsity_to_width(0.25, 226, 11, 4080)])
    np.savez_compressed(HERE / "reference_golden.npz", **out)
    print("wrote", HERE / "reference_golden.npz", sum(v.nbytes for v in out.values()), "bytes raw")


if __name__ == "__main__":
    main()
