"""Ray generation and clipping (reference: nerfs/ray_sampling.py:10-176), stage 1 of the hot path.

Same functions, argument meaning and return shapes as the reference; the arithmetic runs in
csrc/rays.cu.  Rays are packed (…,8) = [o(3), d(3), near, far] fp32."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from .. import ops
from .scene_box import SceneBox


def pack_rays(rays_o: Tensor, rays_d: Tensor, near: Tensor, far: Tensor) -> Tensor:
    return torch.cat([rays_o, rays_d, near, far], dim=-1)


def unpack_rays(rays: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    assert rays.shape[-1] == 8, "packed rays must be (..., 8)"
    flat = rays.view(-1, 8).contiguous()
    return flat[:, :3], flat[:, 3:6], flat[:, 6:7], flat[:, 7:8]


def get_ray_directions(H: int, W: int, fx: float, fy: float, cx: float, cy: float, center_pixels: bool,
                       device: torch.device) -> Tensor:
    """Unit camera-frame (RUB) pinhole directions (H,W,3) (reference :111-136)."""
    return ops.ray_directions(int(H), int(W), float(fx), float(fy), float(cx), float(cy), bool(center_pixels), device)


def get_rays(directions: Tensor, c2w: Tensor, scene_box: Optional[SceneBox] = None, near: Optional[float] = None,
             far: Optional[float] = None, *, aabb_max_bound: float = 1e10, aabb_invalid_value: float = 1e10) -> Tensor:
    """Camera directions + pose -> packed world rays, near/far from the scene box or constants
    (reference :50-108).  (H,W,3) -> (H,W,8); (N,3) -> (N,8)."""
    if directions.ndim == 2 and directions.shape[1] == 3:
        out_shape = (directions.shape[0], 8)
    elif directions.ndim == 3 and directions.shape[-1] == 3:
        out_shape = (directions.shape[0], directions.shape[1], 8)
    else:
        raise ValueError(f"directions must be (H, W, 3) or (N, 3), got {tuple(directions.shape)}")
    if scene_box is None and (near is None or far is None):
        raise ValueError("Provide near/far when scene_box is None")
    aabb = scene_box.aabb if scene_box is not None else None
    rays = ops.get_rays(directions, c2w, aabb, 0.0 if near is None else near, 0.0 if far is None else far,
                        float(aabb_max_bound), float(aabb_invalid_value))
    return rays.view(out_shape)


@torch.no_grad()
def clamp_rays_near_far(rays: Tensor, near_far_override: Optional[Tuple[Optional[float], Optional[float]]], *,
                        eps: float = 1e-6, invalid_value: float = float("inf")) -> Tuple[Tensor, Tensor]:
    """Optional near/far overrides + validity mask (reference :139-176).  With an override tuple
    the rays are cloned and invalid rays get `invalid_value`; with None only the mask is built."""
    if near_far_override is None:
        valid = ops.clamp_near_far_(ops.dev_f32(rays, "rays"), False, None, None, eps, invalid_value)  # read-only
        return rays, valid
    n, f = near_far_override
    out = ops.dev_f32(rays, "rays").clone()
    valid = ops.clamp_near_far_(out, True, n, f, eps, invalid_value)
    return out, valid
