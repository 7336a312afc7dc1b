"""Install the svgb200 operators into an importable copy of the reference (svg-project/Sparse-VideoGen).

The reference's attention processors resolve their operators as module globals
(svg/models/hyvideo/attention.py:12-28); replacing those names is the same mechanism the reference uses
for its own monkey patches (custom_models.py:259-263).  No reference file is edited.

    import svg.models.hyvideo.attention as A
    import svgb200.patch as P
    P.install(A)            # hyvideo / wan / cosmos / cog attention modules
    P.install_kmeans_utils()  # svg.kmeans_utils names used elsewhere
"""
from __future__ import annotations

from . import kmeans_utils as _ku
from . import permute as _pm
from . import placement as _pl

_NAMES = {
    "apply_inverse_permutation_triton": _pm.apply_inverse_permutation_triton,
    "permute_tensor_by_labels_triton": _pm.permute_tensor_by_labels_triton,
    "batch_kmeans_Euclid": _ku.batch_kmeans_Euclid,
    "density_calculation": _ku.density_calculation,
    "dynamic_block_sparse_fwd_flashinfer": _ku.dynamic_block_sparse_fwd_flashinfer,
    "dynamic_block_sparse_fwd_triton": _ku.dynamic_block_sparse_fwd_triton,
    "identify_dynamic_map": _ku.identify_dynamic_map,
    "hunyuan_sparse_head_placement": _pl.hunyuan_sparse_head_placement,
    "hunyuan_hidden_states_placement": _pl.hunyuan_hidden_states_placement,
    "wan_sparse_head_placement": _pl.wan_sparse_head_placement,
    "wan_hidden_states_placement": _pl.wan_hidden_states_placement,
    "cosmos_sparse_head_placement": _pl.cosmos_sparse_head_placement,
    "cosmos_hidden_states_placement": _pl.cosmos_hidden_states_placement,
    "sparse_head_placement": _pl.sparse_head_placement,
    "hidden_states_placement": _pl.hidden_states_placement,
}


def install(module) -> list:
    """Replace every operator name the module already defines.  Returns the names replaced."""
    done = []
    for name, fn in _NAMES.items():
        if hasattr(module, name):
            setattr(module, name, fn)
            done.append(name)
    return done


def install_kmeans_utils(module=None) -> list:
    if module is None:
        import svg.kmeans_utils as module  # noqa: WPS433 (the reference must be importable)
    names = ["batch_kmeans_Euclid", "density_calculation", "dynamic_block_sparse_fwd_flashinfer",
             "dynamic_block_sparse_fwd_triton", "identify_dynamic_map", "euclid_assign_triton",
             "triton_centroid_update_sorted_euclid"]
    for n in names:
        setattr(module, n, getattr(_ku, n))
    return names


def install_native_kernels() -> None:
    """Make `import _kernels` (svg/models/*/attention.py: `sys.path.append("svg/kernels/build/"); import _kernels`)
    resolve to the svgb200 implementation, so the reference's ENABLE_FAST_KERNEL branch runs these kernels."""
    import sys

    from . import _kernels as k

    sys.modules["_kernels"] = k


def install_triton_glue(module) -> list:
    """Replace the Triton glue names a reference module imported (svg/models/wan/custom_models.py:16-17,
    svg/models/wan/attention.py:16)."""
    from . import triton_glue as g

    done = []
    for name in ("triton_rmsnorm_forward", "triton_layernorm_forward", "triton_modulate_shift_forward",
                 "triton_modulate_gate_residual_forward"):
        if hasattr(module, name):
            setattr(module, name, getattr(g, name))
            done.append(name)
    return done
