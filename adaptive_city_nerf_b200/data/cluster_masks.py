"""Offline Voronoi mask generation with per-expert sample boxes on the GPU (reference: scripts/create_clusters.py).

The reference's per-image loop (:799-886) -- pixel directions, world rays clipped to the global scene box, 256 uniform
samples per ray, ray -> expert membership, and (compute_voronoi_opt :386-556, update_aabbs) the streamed per-expert
boxes / sample counts -- as three kernel launches per image (`acn_ray_directions` once per camera model, `acn_get_rays`,
`acn_clamp_near_far`, `acn_route_rays_voronoi` with its box outputs), followed by the reference's post-processing
(:928-962).  Images are independent: with several ranks each takes a strided subset and `distributed.reduce_expert_aabbs`
combines the boxes (the reference's MIN / MAX / SUM all-reduces, :928-932).

Only the per-ray work is here; reading metadata files, saving masks and the statistics text stay the reference's Python."""
from __future__ import annotations

from typing import Callable, Dict, Iterable, Optional, Sequence, Tuple

import torch
from torch import Tensor

from .. import ops
from ..nerfs.ray_sampling import clamp_rays_near_far, get_ray_directions, get_rays
from ..nerfs.scene_box import SceneBox


def new_expert_boxes(K: int, device) -> Tuple[Tensor, Tensor, Tensor]:
    """(mins, maxs, counts) in their streaming start state (scripts/create_clusters.py:792-794)."""
    return (torch.full((K, 3), float("inf"), dtype=torch.float32, device=device),
            torch.full((K, 3), float("-inf"), dtype=torch.float32, device=device),
            torch.zeros(K, dtype=torch.int64, device=device))


@torch.no_grad()
def voronoi_masks_and_boxes(images: Iterable[Dict], centroids: Tensor, aabb_global: Tensor, *, ray_samples: int = 256,
                            boundary_margin: float = 1.1, cluster_2d: bool = True, center_pixels: bool = True,
                            near_far_override: Tuple[Optional[float], Optional[float]] = (None, None),
                            boxes: Optional[Tuple[Tensor, Tensor, Tensor]] = None,
                            on_mask: Optional[Callable[[int, Tensor, Tensor], None]] = None,
                            device=None) -> Tuple[Tensor, Tensor, Tensor]:
    """images: dicts with H, W, intrinsics (fx, fy, cx, cy) and c2w (3,4) as the reference's metadata files hold them.
    on_mask(i, mask (H,W,K) bool -- already ANDed with the ray validity, valid (H,W) bool) receives every image's masks
    (scripts/create_clusters.py:868-879).  -> the streamed (mins, maxs, counts), un-finalised (`finalize_expert_boxes`)."""
    device = torch.device(device) if device is not None else centroids.device
    cen = centroids.to(device=device, dtype=torch.float32).contiguous()
    K = cen.shape[0]
    box = SceneBox(aabb=aabb_global.to(device=device, dtype=torch.float32))
    mins, maxs, counts = boxes if boxes is not None else new_expert_boxes(K, device)
    dims = 2 if cluster_2d else 3
    dirs_cache: Dict[tuple, Tensor] = {}
    for i, md in enumerate(images):
        H, W = int(md["H"]), int(md["W"])
        fx, fy, cx, cy = (float(v) for v in md["intrinsics"])
        key = (H, W, fx, fy, cx, cy)
        dirs = dirs_cache.get(key)
        if dirs is None:
            dirs_cache.clear()
            dirs = dirs_cache[key] = get_ray_directions(H, W, fx, fy, cx, cy, center_pixels, device)
        c2w = torch.as_tensor(md["c2w"], dtype=torch.float32).to(device)
        rays = get_rays(dirs, c2w, scene_box=box, aabb_max_bound=1e10, aabb_invalid_value=float("inf")).view(-1, 8)
        rays, valid = clamp_rays_near_far(rays, near_far_override)
        mask = ops.route_rays_voronoi(rays, int(ray_samples), cen, dims, float(boundary_margin), aabb_out=(mins, maxs, counts))
        if on_mask is not None:
            on_mask(i, (mask & valid[:, None]).view(H, W, K), valid.view(H, W))
    return mins, maxs, counts


@torch.no_grad()
def finalize_expert_boxes(mins: Tensor, maxs: Tensor, counts: Tensor, centroids: Tensor, aabb_global: Tensor,
                          box_margin: float = 0.0, pose_scale: float = 1.0) -> Tuple[Tensor, Tensor]:
    """scripts/create_clusters.py:934-962: clamp to the global box, give empty experts an epsilon box around their
    centroid, optional dilation, and the global altitude (x) range for everyone.  Call after the cross-rank reduction."""
    dev = mins.device
    lo, hi = aabb_global[0].to(dev).float(), aabb_global[1].to(dev).float()
    mins, maxs = torch.maximum(mins, lo), torch.minimum(maxs, hi)
    empties = counts == 0
    if bool(empties.any()):
        eps = torch.clamp((hi - lo).abs() * 1e-6, min=1e-7)
        cc = torch.minimum(torch.maximum(centroids.to(dev).float(), lo), hi)
        mins[empties] = torch.maximum(cc[empties] - eps, lo)
        maxs[empties] = torch.minimum(cc[empties] + eps, hi)
    if box_margin and box_margin > 0.0:
        m = float(box_margin) / float(pose_scale)
        mins, maxs = torch.maximum(mins - m, lo), torch.minimum(maxs + m, hi)
    mins[:, 0], maxs[:, 0] = lo[0], hi[0]
    return mins, maxs
