"""`RamRaysDataset` with the per-image ray work on the GPU (reference: data/ram_rays_dataset.py:46-258).

The reference spends its constructor in a pool of CPU processes: per image, pixel directions, camera-to-world rays
clipped to the scene box, the keep-mask, near / far clamping and validity filtering -- ~15 whole-image torch ops on one
core.  Here each image is three kernel launches (`acn_ray_directions` once per camera model, `acn_get_rays`,
`acn_clamp_near_far`) plus a boolean compaction on the device; what stays on the host is what has to (reading the image
and mask files through the metadata object's own `load_image` / `load_mask`).

Same constructor arguments, attributes (`_rgbs (M,3)` fp32 in [0,1], `_rays (M,8)`, `_img_indices (M,)` int32,
`_num_images`, `_img_unique_ids`) and `Dataset` interface; the tensors live on `device` (default: the scene box's
device, else cuda:0), so `TaskGrid` / training batches index them without a host round trip."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
from torch.utils.data import Dataset

from ..nerfs.ray_sampling import clamp_rays_near_far, get_ray_directions, get_rays


def _val_balancing(keep_mask: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """data/ram_rays_dataset.py:236-258 (Mega-NeRF): drop the right half, and re-admit as many currently masked-out pixels
    of the left half as the right half had kept ones (a random choice of them)."""
    keep_mask = keep_mask.view(H, W).clone()
    left, right = keep_mask[:, : W // 2], keep_mask[:, W // 2:]
    discard_pos = int(right.sum().item())
    if discard_pos > 0:
        candidates = torch.arange(H * W, device=keep_mask.device).view(H, W)[:, : W // 2]
        not_kept_left = candidates[~left]
        if not_kept_left.numel() > 0:
            perm = torch.randperm(not_kept_left.numel(), device=keep_mask.device)
            flat = keep_mask.view(-1)
            flat[not_kept_left[perm[:discard_pos]]] = True
            keep_mask = flat.view(H, W)
    keep_mask[:, W // 2:] = False
    return keep_mask.view(-1).bool()


class RamRaysDataset(Dataset):
    def __init__(self, metadata_items: List, center_pixels: bool, val_balancing: bool = False,
                 ray_gen_kwargs: Optional[dict] = None, num_workers: Optional[int] = None, device=None):
        """`num_workers` is accepted for call compatibility and ignored: there is no process pool to size."""
        super().__init__()
        if ray_gen_kwargs is None or "scene_box" not in ray_gen_kwargs:
            raise ValueError("ray_gen_kwargs must contain keys: 'scene_box' and 'near_far_override'")
        scene_box = ray_gen_kwargs["scene_box"]
        near_far_override = ray_gen_kwargs.get("near_far_override", None)
        if device is None:
            device = scene_box.aabb.device if scene_box.aabb.is_cuda else torch.device("cuda", torch.cuda.current_device())
        device = torch.device(device)
        scene_box = scene_box.to(device)
        rgbs, rays, indices = [], [], []
        dirs_key, dirs = None, None
        with torch.no_grad():
            for md in metadata_items:
                if md is None:
                    continue
                img = md.load_image()
                if img is None:
                    continue
                H, W = int(md.H), int(md.W)
                if img.ndim == 2 and img.shape[-1] == 3:
                    img = img.view(H, W, 3)
                elif img.ndim == 3 and img.shape[0] == 3:
                    img = img.permute(1, 2, 0).contiguous()
                elif not (img.ndim == 3 and img.shape[-1] == 3):
                    continue
                img = img.to(device, non_blocking=True)
                keep = md.load_mask()
                if keep is not None:
                    keep = keep.to(device).view(H, W)
                if getattr(md, "is_val", False) and val_balancing:
                    if keep is None:
                        keep = torch.ones(H, W, dtype=torch.bool, device=device)
                    keep = _val_balancing(keep, H, W)
                if keep is not None and int(keep.sum().item()) == 0:
                    continue
                fx, fy, cx, cy = (float(v) for v in md.intrinsics)
                key = (H, W, fx, fy, cx, cy)
                if key != dirs_key:
                    dirs_key, dirs = key, get_ray_directions(H, W, fx, fy, cx, cy, center_pixels, device=device)
                image_rays = get_rays(dirs, torch.as_tensor(md.c2w, dtype=torch.float32).to(device), scene_box=scene_box).view(-1, 8)
                img = img.view(-1, 3)
                if keep is not None:
                    flat = keep.view(-1)
                    image_rays, img = image_rays[flat], img[flat]
                image_rays, valid = clamp_rays_near_far(image_rays, near_far_override=near_far_override)
                if not bool(valid.any()):
                    continue
                image_rays = image_rays[valid]
                # a TENSOR divisor: torch's CUDA div-by-Python-scalar multiplies by the rounded reciprocal (1 ulp off the
                # reference's CPU `/ 255.0`, data/ram_rays_dataset.py:211); tensor / tensor is an IEEE division
                img = img[valid].to(torch.float32) / torch.full((), 255.0, dtype=torch.float32, device=device)
                rgbs.append(img.contiguous())
                rays.append(image_rays.contiguous())
                indices.append(torch.full((img.shape[0],), int(md.image_index), dtype=torch.int32, device=device))
        if not rgbs:
            self._rgbs = torch.zeros((0, 3), dtype=torch.float32, device=device)
            self._rays = torch.zeros((0, 8), dtype=torch.float32, device=device)
            self._img_indices = torch.zeros((0,), dtype=torch.int32, device=device)
            self._num_images = 0
            self._img_unique_ids = []
        else:
            self._rgbs = torch.cat(rgbs, dim=0).contiguous()
            self._rays = torch.cat(rays, dim=0).contiguous()
            self._img_indices = torch.cat(indices, dim=0).contiguous()
            self._num_images = len(rgbs)
            self._img_unique_ids = torch.unique(self._img_indices).cpu().tolist()

    def __len__(self) -> int:
        return self._rgbs.shape[0]

    def __getitem__(self, idx) -> Dict[str, torch.Tensor]:
        return {"rgbs": self._rgbs[idx], "rays": self._rays[idx], "img_indices": self._img_indices[idx]}

    _apply_meganerf_val_balancing_static = staticmethod(_val_balancing)
