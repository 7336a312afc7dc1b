"""Golden vectors for the SVG1 profiling path, the HunyuanVideo mask_mod and the ops-API test oracles, produced by
EXECUTING THE REAL REFERENCE on CPU (build container; /root/reference mounted or baseline/_ref installed).

    python tests/golden/make_golden_svg1.py      ->  tests/golden/svg1_golden.npz

The reference modules import `diffusers`, `cuvs`, `matplotlib` at module scope; tests/golden/ref_import.py
supplies inert stubs for exactly those names (none is touched by the functions run here).  `get_attention_mask`
ends in `.cuda()` (hyvideo/utils.py:92, wan/utils.py:109, cog/utils.py:63): `torch.Tensor.cuda` is made the identity
while it runs, nothing else is altered.

Functions executed (reference file:line):
  svg/models/hyvideo/utils.py:20-44   generate_temporal_head_mask_mod
  svg/models/hyvideo/utils.py:47-93   get_attention_mask          (wan/utils.py:63-110, cog/utils.py:61-88)
  svg/models/hyvideo/utils.py:142-151 sparsity_to_width
  svg/models/hyvideo/attention.py:375-399  Hunyuan_SVGAttn_Processor2_0.sample_mse   (bf16, as the live path)
  svg/models/wan/attention.py:211-233      WanAttn_SVGAttn_Processor2_0.sample_mse
  svg/models/cog/attention.py:118-145      CogVideoX_SparseAttn_Processor2_0.sample_mse
  svg/models/hyvideo/attention.py:657-702  Hunyuan_SAPAttn_Processor2_0.dynamic_map_post_processing
  svg/kernels/test/test_sparse_attn.py:20-157  ref_gen_temporal_mask, ref_gen_spatial_mask, gen_mask_block2element,
                                               ref_torch_attn_impl
  svg/kernels/ops/attention_ops.py:9-104       _gen_temporal_mask, _gen_spatial_mask (BSR)
  svg/kernels/ops/attention_ops_wan.py:48-129  gen_temporal_mask (BSR), ref_gen_temporal_mask
  svg/models/wan/utils.py:130-185              gen_temporal_mask (adds the first-frame region)
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import ref_import as R  # noqa: E402
from gen_inputs import checksum, smse_inputs  # noqa: E402


class _cuda_is_identity:
    def __enter__(self):
        self.orig = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self

    def __exit__(self, *a):
        torch.Tensor.cuda = self.orig


def bf16_bits(t):
    return t.contiguous().view(torch.int16).numpy()


def main():
    out = {}
    hy_u = R.import_model_module("hyvideo", "utils")
    wan_u = R.import_model_module("wan", "utils")
    cog_u = R.import_model_module("cog", "utils")
    hy_a = R.import_model_module("hyvideo", "attention")
    wan_a = R.import_model_module("wan", "attention")
    cog_a = R.import_model_module("cog", "attention")

    # ---- HY executed mask_mod on three parameter sets (q, kv over the whole small sequence)
    for i, (ctx, plen, F, P, mul) in enumerate([(40, 17, 4, 150, 1.3), (64, 64, 3, 200, 0.7), (16, 0, 5, 133, 2.2)]):
        S = ctx + F * P
        mod = hy_u.generate_temporal_head_mask_mod(ctx, plen, F, P, mul)
        qi = torch.arange(S).view(-1, 1)
        ki = torch.arange(S).view(1, -1)
        out[f"hy_mm{i}_params"] = np.array([ctx, plen, F, P, mul], dtype=np.float64)
        out[f"hy_mm{i}"] = np.packbits(mod(None, None, qi, ki).numpy())
    out["hy_s2w"] = np.array([hy_u.sparsity_to_width(0.25, 256, 33, 3600), hy_u.sparsity_to_width(0.30, 256, 33, 3600),
                              hy_u.sparsity_to_width(0.4, 64, 5, 1000)])

    # ---- profiling masks (get_attention_mask) of the three model families, small sizes
    with _cuda_is_identity():
        for name, (ctx, F, P, max_row) in {"hy": (16, 3, 200, 500), "hy2": (40, 4, 390, 10000)}.items():
            for mn in ("spatial", "temporal"):
                m = hy_u.get_attention_mask(mn, max_row, ctx, F, P, device="cpu")
                out[f"prof_{name}_{mn}"] = np.packbits(m.bool().numpy())
                out[f"prof_{name}_dims"] = np.array([ctx, F, P, max_row, m.shape[0], m.shape[1]])
        for name, (F, P, max_row) in {"wan": (4, 150, 10000), "wan2": (3, 260, 400)}.items():
            for mn in ("spatial", "temporal"):
                m = wan_u.get_attention_mask(mn, max_row, 0, F, P)
                out[f"prof_{name}_{mn}"] = np.packbits(m.bool().numpy())
                out[f"prof_{name}_dims"] = np.array([0, F, P, max_row, m.shape[0], m.shape[1]])
        for name, (ctx, F, P) in {"cog": (30, 3, 200), "cog2": (226, 2, 390)}.items():
            for mn in ("spatial", "temporal"):
                m = cog_u.get_attention_mask(mn, ctx, F, P)
                out[f"prof_{name}_{mn}"] = np.packbits(m.bool().numpy())
                out[f"prof_{name}_dims"] = np.array([ctx, F, P, m.shape[0], m.shape[0], m.shape[1]])

        # ---- sample_mse of the real processors, bf16 like the live path; rows come from the CPU generator (:381)
        def run_smse(cls, masks, seed, cfg, H, S, D, nrows, max_row=None):
            q, k, v = smse_inputs(seed, cfg, H, S, D)
            proc = cls.__new__(cls)
            proc.layer_idx = 0
            cls.attention_masks = masks
            cls.num_sampled_rows = nrows
            if max_row is not None:
                cls.sample_mse_max_row = max_row
            torch.manual_seed(seed + 100)
            state = torch.get_rng_state()
            mses = proc.sample_mse(q, k, v)
            torch.set_rng_state(state)
            hi = max_row if max_row is not None else S
            rows = torch.randint(low=0, high=hi, size=(min(nrows, S),))
            return q, k, v, rows, mses

        cases = {
            "hy": (hy_a.Hunyuan_SVGAttn_Processor2_0, lambda: [hy_u.get_attention_mask(n, 500, 16, 3, 200, device="cpu")
                                                               for n in ("spatial", "temporal")], (1, 3, 616, 64, 24, 500)),
            "wan": (wan_a.WanAttn_SVGAttn_Processor2_0, lambda: [wan_u.get_attention_mask(n, 10000, 0, 4, 150)
                                                             for n in ("spatial", "temporal")], (1, 3, 600, 64, 24, 600)),
            "cog": (cog_a.CogVideoX_SparseAttn_Processor2_0, lambda: [cog_u.get_attention_mask(n, 30, 3, 200)
                                                                      for n in ("spatial", "temporal")], (1, 3, 630, 64, 24, None)),
        }
        for name, (cls, mk, (cfg, H, S, D, nrows, max_row)) in cases.items():
            for rep in range(2 if name != "cog" else 4):
                q, k, v, rows, mses = run_smse(cls, mk(), 10 + rep, cfg, H, S, D, nrows, max_row)
                out[f"smse_{name}{rep}_in"] = np.array([10 + rep, cfg, H, S, D, checksum(q, k, v)], dtype=np.float64)
                out[f"smse_{name}{rep}_rows"] = rows.numpy()
                out[f"smse_{name}{rep}_mses"] = mses.float().numpy()
                out[f"smse_{name}{rep}_best"] = torch.argmin(mses, dim=0).numpy()

    # ---- HunyuanVideo SVG2 prompt / padding blocks (dynamic_map_post_processing)
    g = torch.Generator().manual_seed(3)
    H, V, ctx, plen, QC, KC, D = 2, 40, 8, 5, 3, 4, 4
    proc = hy_a.Hunyuan_SAPAttn_Processor2_0.__new__(hy_a.Hunyuan_SAPAttn_Processor2_0)
    q, k, v = (torch.randn(1, H, V + ctx, D, generator=g) for _ in range(3))
    qp, kp, vp = (torch.randn(1, H, V, D, generator=g) for _ in range(3))
    dyn = torch.rand(1, H, QC, KC, generator=g) > 0.5
    qsz = torch.tensor([[[10, 20, 10], [40, 0, 0]]], dtype=torch.int32)
    ksz = torch.tensor([[[10, 10, 10, 10], [0, 40, 0, 0]]], dtype=torch.int32)
    qidx = torch.stack([torch.randperm(V, generator=g) for _ in range(H)]).int()
    r = proc.dynamic_map_post_processing(qp, kp, vp, q.clone(), k.clone(), v.clone(), dyn, qsz, ksz, qidx, V, ctx, plen,
                                         ctx - plen)
    out.update(pp_dims=np.array([H, V, ctx, plen, QC, KC, D]), pp_q=q.numpy(), pp_qp=qp.numpy(), pp_dyn=dyn.numpy(),
               pp_qsz=qsz.numpy(), pp_ksz=ksz.numpy(), pp_qidx=qidx.numpy(), pp_out_q=r[0].numpy(), pp_out_dyn=r[3].numpy(),
               pp_out_qsz=r[4].numpy(), pp_out_ksz=r[5].numpy(), pp_out_qidx=r[6].numpy())

    # ---- ops API: the reference TEST oracles and the BSR generators of the ops modules (they end in .cuda(); the Wan
    # generator also plots the mask through matplotlib -- that side effect is switched off)
    t = R.import_kernel_test("test_sparse_attn")
    ops = R.import_ops("attention_ops")
    ops_wan = R.import_ops("attention_ops_wan")
    ops_wan.visualize_attention_mask = lambda *a, **k: None
    F, P, mul = 6, 40, 1.7
    with _cuda_is_identity():
        bm = t.ref_gen_temporal_mask(F, P, mul)
        out["ops_ref_temporal"] = bm
        bm2 = t.ref_gen_spatial_mask(F, P, 1)
        out["ops_ref_spatial"] = bm2
        em = t.gen_mask_block2element(bm, (P // 10, P // 10), 7)
        out["ops_b2e"] = np.packbits(em.numpy())
        out["ops_b2e_shape"] = np.array(em.shape)
        g = torch.Generator().manual_seed(4)
        S = 7 + F * P
        q, k, v = (torch.randn(S, 2, 16, generator=g) for _ in range(3))
        o = t.ref_torch_attn_impl(q, k, v, em)
        out.update(ops_attn_q=q.numpy(), ops_attn_k=k.numpy(), ops_attn_v=v.numpy(), ops_attn_o=o.float().numpy())
        import contextlib
        import io

        with contextlib.redirect_stdout(io.StringIO()):
            ip, ix, shape = ops._gen_temporal_mask(F, P, mul)
            out.update(ops_bsr_t_indptr=ip.numpy(), ops_bsr_t_indices=ix.numpy(), ops_bsr_t_shape=np.array(shape))
            ip, ix, shape = ops._gen_spatial_mask(F, P, 1)
            out.update(ops_bsr_s_indptr=ip.numpy(), ops_bsr_s_indices=ix.numpy(), ops_bsr_s_shape=np.array(shape))
            Fw, Pw, mw = 5, 480, 1.3
            ip, ix, shape = ops_wan.gen_temporal_mask(Fw, Pw, mw)
            out.update(opsw_bsr_indptr=ip.numpy(), opsw_bsr_indices=ix.numpy(), opsw_bsr_shape=np.array(shape))
            out["opsw_ref"] = ops_wan.ref_gen_temporal_mask(Fw, Pw, mw)
            ip, ix, shape = wan_u.gen_temporal_mask(Fw, Pw, mw)
            out.update(wanu_bsr_indptr=ip.numpy(), wanu_bsr_indices=ix.numpy(), wanu_bsr_shape=np.array(shape))
    out["opsw_params"] = np.array([Fw, Pw, mw])
    out["ops_params"] = np.array([F, P, mul])

    np.savez_compressed(HERE / "svg1_golden.npz", **out)
    print("wrote", HERE / "svg1_golden.npz", sum(v.nbytes for v in out.values()) // 1024, "KiB raw")


if __name__ == "__main__":
    main()
