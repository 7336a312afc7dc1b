"""Golden vectors for the pre-attention chain from the REFERENCE'S OWN ground-truth functions.

The reference's tests for its native `_kernels` ops (svg/kernels/test/test_{rms_norm,layer_norm,apply_rope_txtlast,
apply_rope_complex}.py) compare the CUDA kernels with small host functions (`ref_host_*`, `replica_host_rms_norm`).
Those test modules import `_kernels` at the top (not buildable here), so this script extracts just the function
definitions with `ast`, executes them unmodified, and stores their outputs.  Runs only in the build container:

    python tests/golden/make_golden_prep.py      ->  tests/golden/prep_golden.npz
"""
import ast
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
REF_TESTS = Path("/root/reference/svg/kernels/test")


def extract(fname, names):
    src = (REF_TESTS / fname).read_text()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(keep) == len(names), (fname, names)
    ns = {"torch": torch, "Tuple": tuple}
    from typing import Tuple
    ns["Tuple"] = Tuple
    exec(compile(ast.Module(body=keep, type_ignores=[]), str(REF_TESTS / fname), "exec"), ns)
    return [ns[n] for n in names]


def f32(t):
    return t.float().numpy()


def main():
    g = torch.Generator().manual_seed(0)
    out = {}
    ref_rms, replica_rms = extract("test_rms_norm.py", ["ref_host_rms_norm", "replica_host_rms_norm"])
    (ref_ln,) = extract("test_layer_norm.py", ["ref_host_layer_norm"])
    (ref_rope,) = extract("test_apply_rope_txtlast.py", ["ref_host_apply_rope"])
    (ref_rope_c,) = extract("test_apply_rope_complex.py", ["ref_host_apply_rope_complex"])

    for n in (32, 64, 128, 256):
        x = torch.randn(23, n, generator=g).bfloat16()
        gm = torch.randn(n, generator=g).bfloat16()
        bt = torch.randn(n, generator=g).bfloat16()
        out[f"rms_x_{n}"], out[f"rms_g_{n}"], out[f"ln_b_{n}"] = f32(x), f32(gm), f32(bt)
        out[f"rms_ref_{n}"] = f32(ref_rms(x, gm).to(x))
        out[f"rms_replica_{n}"] = f32(replica_rms(x, gm).to(x))
        out[f"ln_ref_{n}"] = f32(ref_ln(x, gm, bt))

    for D in (64, 128):
        B, H, S, T = 1, 2, 83, 19
        q = torch.randn(B, H, S, D, generator=g).bfloat16()
        cos = torch.randn(S - T, D, generator=g)
        sin = torch.randn(S - T, D, generator=g)
        out[f"rope_q_{D}"], out[f"rope_cos_{D}"], out[f"rope_sin_{D}"] = f32(q), cos.numpy(), sin.numpy()
        out[f"rope_txtlast_{D}"] = f32(ref_rope(q[:, :, :-T, :], cos, sin))
        out[f"rope_txtfirst_{D}"] = f32(ref_rope(q[:, :, T:, :], cos, sin))
        qh = q.half()
        fr = torch.complex(torch.randn(S - T, D // 2, generator=g), torch.randn(S - T, D // 2, generator=g))
        out[f"ropec_re_{D}"], out[f"ropec_im_{D}"] = fr.real.numpy().copy(), fr.imag.numpy().copy()
        out[f"ropec_out_{D}"] = f32(ref_rope_c(qh[:, :, T:, :], fr))
    np.savez_compressed(HERE / "prep_golden.npz", **out)
    print("wrote", HERE / "prep_golden.npz", {k: v.shape for k, v in list(out.items())[:4]})


if __name__ == "__main__":
    main()
