"""Oracle: token permutation and SVG1 head placement.  TEST INFRASTRUCTURE ONLY.

Restates (relative to /root/reference):
  permute_tensor_by_labels / apply_inverse_permutation   svg/kmeans_utils.py:820-849
  _permute_kernel / _inverse_permute_kernel              svg/kernels/triton/permute.py:12-75
  ref_hunyuan_sparse_head_placement                      svg/models/hyvideo/placement.py:156-184
  ref_hunyuan_hidden_states_placement                    svg/models/hyvideo/placement.py:390-401
  cog text-first variants                                svg/models/cog/placement.py
"""
from __future__ import annotations

import numpy as np
import torch


def stable_argsort_labels(labels: np.ndarray) -> np.ndarray:
    """sorted_indices = argsort(labels) with ties in ascending token order.  The reference calls
    torch.argsort (permute.py:113), which is unstable; the engine defines the stable order and the
    attention result is invariant to the intra-cluster order (SURVEY §7 hard part (a))."""
    return np.argsort(labels, axis=-1, kind="stable").astype(np.int32)


def permute_gather(x: torch.Tensor, perm: np.ndarray) -> torch.Tensor:
    """Y[bh, s, :] = X[bh, perm[bh, s], :]   (permute.py:12-43).  x: [BH,S,D]."""
    idx = torch.from_numpy(perm.astype(np.int64))
    return torch.gather(x, 1, idx.unsqueeze(-1).expand(-1, -1, x.shape[-1]))


def permute_scatter(x: torch.Tensor, perm: np.ndarray) -> torch.Tensor:
    """Y[bh, perm[bh, s], :] = X[bh, s, :]   (permute.py:46-75)."""
    idx = torch.from_numpy(perm.astype(np.int64))
    out = torch.empty_like(x)
    out.scatter_(1, idx.unsqueeze(-1).expand(-1, -1, x.shape[-1]), x)
    return out


def _to_token_major(t, F, P):
    """hunyuan_token_reorder_to_token_major (placement.py:6-18): video rows [F,P] -> [P,F]."""
    BH, V, D = t.shape
    return t.reshape(BH, F, P, D).transpose(1, 2).reshape(BH, V, D)


def _to_frame_major(t, F, P):
    """hunyuan_token_reorder_to_frame_major (placement.py:21-32): video rows [P,F] -> [F,P]."""
    BH, V, D = t.shape
    return t.reshape(BH, P, F, D).transpose(1, 2).reshape(BH, V, D)


def head_placement(x: torch.Tensor, best_mask_idx, ctx: int, F: int, P: int, text_first: bool = False,
                   inverse: bool = False) -> torch.Tensor:
    """x: [BH,S,D].  Heads with best_mask_idx == 1 get their video part reordered (forward:
    frame-major -> token-major, placement.py:156-184; inverse: placement.py:390-401); text rows
    (last `ctx` rows for HY, first `ctx` for Cog) and all other heads are copied."""
    out = x.clone()
    sel = torch.as_tensor(np.asarray(best_mask_idx) == 1)
    if sel.any():
        V = F * P
        lo = ctx if text_first else 0
        vid = x[sel][:, lo:lo + V]
        out[sel, lo:lo + V] = _to_frame_major(vid, F, P) if inverse else _to_token_major(vid, F, P)
    return out
