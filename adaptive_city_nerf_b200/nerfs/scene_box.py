"""Axis-aligned scene box (reference: nerfs/scene_box.py:11-107).

Only the pieces the rendering hot path touches are mirrored: the (2,3) `aabb` tensor with
min/max/center/extent accessors, `.to`, `within`, `get_diagonal_length` and the slab test
`ray_aabb_intersect`, which runs in csrc/rays.cu with the reference's exact rounding sequence
(eps-guarded IEEE reciprocal, then (bound - o) * inv)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import torch
from torch import Tensor

from .. import ops


@dataclass
class SceneBox:
    aabb: Tensor  # (2,3): [min, max]

    @property
    def min(self) -> Tensor:
        return self.aabb[0]

    @property
    def max(self) -> Tensor:
        return self.aabb[1]

    @property
    def center(self) -> Tensor:
        return (self.aabb[0] + self.aabb[1]) * 0.5

    @property
    def extent(self) -> Tensor:
        return self.aabb[1] - self.aabb[0]

    def to(self, dev) -> "SceneBox":
        dev = torch.device(dev) if isinstance(dev, str) else dev
        return SceneBox(aabb=self.aabb.to(dev))

    def __repr__(self) -> str:
        mn = ", ".join(f"{v:.3f}" for v in self.min.cpu().tolist())
        mx = ", ".join(f"{v:.3f}" for v in self.max.cpu().tolist())
        return f"SceneBox (min=[{mn}], max=[{mx}], diag={self.get_diagonal_length().item():.3f})"

    def ray_aabb_intersect(self, origins: Tensor, directions: Tensor, eps: float = 1e-8,
                           max_bound: float = 1e10, invalid_value: float = 1e10) -> Tuple[Tensor, Tensor]:
        """Slab test with clamping to [0, max_bound]; rays that miss get `invalid_value` in both
        outputs (reference scene_box.py:45-107).  origins/directions (N,3) CUDA tensors."""
        assert self.aabb.shape == (2, 3), "aabb must be (2,3)"
        return ops.aabb_intersect(origins, directions, self.aabb, eps, max_bound, invalid_value)

    def within(self, pts: Tensor, inclusive: bool = False) -> Tensor:
        if inclusive:
            return (pts >= self.aabb[0]).all(dim=-1) & (pts <= self.aabb[1]).all(dim=-1)
        return (pts > self.aabb[0]).all(dim=-1) & (pts < self.aabb[1]).all(dim=-1)

    def get_diagonal_length(self) -> Tensor:
        return torch.linalg.norm(self.aabb[1] - self.aabb[0])
