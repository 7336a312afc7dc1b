"""Oracle: flash-k-means, dynamic map, density.  TEST INFRASTRUCTURE ONLY.

Restates (relative to /root/reference/svg/kmeans_utils.py):
  density_calculation                   :13-31
  _euclid_assign_kernel                 :464-554   (torch formulation kept by the reference at :631-635)
  triton_centroid_update_sorted_euclid  :375-421
  _euclid_iter / batch_kmeans_Euclid    :629-643 / :684-733
  weighted_softmax / identify_dynamic_map  :852-896

Parity status: density / identify_dynamic_map are checked against the reference functions run on
CPU (tests/golden/make_golden.py).  The k-means kernels are Triton (GPU only) and untested by the
reference -> restated from source, parity UNPINNED by reference outputs.
"""
from __future__ import annotations

import torch


def density_calculation(dynamic_map, q_cluster_sizes, k_cluster_sizes):
    """:13-31.  [cfg,H,QC,KC] bool, [cfg,H,QC], [cfg,H,KC] -> [cfg,H] (exact rational in float64)."""
    blk = q_cluster_sizes[..., :, None].double() * k_cluster_sizes[..., None, :].double()
    return ((blk * dynamic_map.double()).sum(dim=(2, 3)) / blk.sum(dim=(2, 3))).float()


def row_sqnorm(x: torch.Tensor) -> torch.Tensor:
    """x_sq = (x**2).sum(-1) in the input dtype (:704): squares rounded to bf16/fp16, summed with
    fp32 accumulation, result rounded to the input dtype; returned as fp32."""
    return (x ** 2).sum(dim=-1).float()


def euclid_assign(x: torch.Tensor, c: torch.Tensor, x_sq: torch.Tensor) -> torch.Tensor:
    """:464-554.  x [B,N,D], c [B,K,D] (16-bit), x_sq fp32 [B,N] -> labels int64 [B,N] plus the
    (best, second-best) distance margin used by tests to skip numerically ambiguous points.
    dist = max(0, x_sq + sum_d(round16(c*c)) - 2 * (x . c)),  first minimum wins."""
    cent_sq = (c * c).float().sum(dim=-1)  # fp32 sum of 16-bit-rounded squares (:531)
    cross = torch.einsum("bnd,bkd->bnk", x.float(), c.float())
    dist = (x_sq[:, :, None] + cent_sq[:, None, :] - 2.0 * cross).clamp_min(0.0)
    labels = dist.argmin(dim=-1)
    top2 = torch.topk(dist, k=min(2, dist.shape[-1]), dim=-1, largest=False).values
    margin = (top2[..., -1] - top2[..., 0]) if dist.shape[-1] > 1 else torch.full_like(top2[..., 0], 1e30)
    return labels, margin


def centroid_update(x: torch.Tensor, labels: torch.Tensor, old_c: torch.Tensor):
    """:375-421.  fp32 mean of members; empty clusters keep the old centroid; cast to x dtype."""
    B, N, D = x.shape
    K = old_c.shape[1]
    sums = torch.zeros(B, K, D, dtype=torch.float32)
    sums.scatter_add_(1, labels[:, :, None].expand(-1, -1, D), x.float())
    cnts = torch.zeros(B, K, dtype=torch.int64)
    cnts.scatter_add_(1, labels, torch.ones_like(labels))
    mean = sums / cnts.clamp(min=1).unsqueeze(-1).float()
    new_c = torch.where((cnts == 0).unsqueeze(-1), old_c.float(), mean)
    return new_c.to(x.dtype), cnts.to(torch.int32)


def batch_kmeans_euclid(x, n_clusters, max_iters, tol=1e-4, init_centroids=None):
    """:684-733.  Returns (labels, centroids, sizes, n_iter) with the reference's exact control
    flow: labels/sizes come from the last assignment; centroids are the updated ones unless the loop
    broke on `shift < tol`, in which case they are the ones the last assignment was made against."""
    assert init_centroids is not None, "random init (:708) uses the GPU generator; pass centroids"
    x_sq = row_sqnorm(x)
    c = init_centroids
    it = 0
    for it in range(max_iters):
        labels, _ = euclid_assign(x, c, x_sq)
        c_new, sizes = centroid_update(x, labels, c)
        shift = (c_new.float() - c.float()).norm(dim=-1).max()
        if shift < tol:
            break
        c = c_new
    return labels, c, sizes, it + 1


def weighted_softmax(scores, weights):
    """:852-861."""
    dt = scores.dtype
    s = scores.float()
    w = weights.float()
    e = torch.exp(s - s.max(dim=-1, keepdim=True)[0])
    we = w * e
    return (we / we.sum(dim=-1, keepdim=True).clamp(min=1e-12)).to(dt)


def identify_dynamic_map(qc, kc, q_sizes, k_sizes, p, min_kc_ratio=0.0):
    """:864-896 with the tie rule the engine defines (stable descending sort: the lower column
    index first).  qc [B,H,QC,D], kc [B,H,KC,D] in the model dtype (bf16): scores, probabilities and
    their running sum are all rounded to that dtype exactly as the reference's torch ops do."""
    B, H, QC, D = qc.shape
    KC = kc.shape[2]
    scores = torch.matmul(qc, kc.transpose(-2, -1)) / (D ** 0.5)
    probs = weighted_softmax(scores, k_sizes.unsqueeze(-2).float())
    sorted_probs, sorted_idx = torch.sort(probs, dim=-1, descending=True, stable=True)
    cums = torch.cumsum(sorted_probs, dim=-1)
    remove = cums > p
    remove[..., 1:] = remove[..., :-1].clone()
    remove[..., 0] = False
    if min_kc_ratio > 0:
        remove[..., : int(min_kc_ratio * KC)] = False
    out = torch.zeros(B, H, QC, KC, dtype=torch.bool)
    out.scatter_(-1, sorted_idx, ~remove)
    return out
