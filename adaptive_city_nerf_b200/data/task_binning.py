"""Task-grid binning of an expert's training rays.

Reference: data/task_dataset.py `TaskDataset.__init__` -> `_init_region_aabb` (:230-237), `_build_cell_bounds`
(:174-197) and `_route_and_bin` (:544-627) with `routing_policy="dda"`, the policy nerf_runner.py:201-209 passes.  The
reference walks the grid with ~64 x 25 elementwise launches over all rays plus an argsort and a Python loop over the
cells; here the per-ray work is ONE kernel (`acn_dda_route_rays`) and the bins come from the bucketing kernel of the
expert dispatch.  The handful of tiny host-side tensors (region box, cell bounds, tolerances) are built with the same
torch ops as the reference so they are bit-identical."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from .. import ops
from .._lib import check, ctx, dev_f32, lib, ptr, stream


class TaskGrid:
    """Region box + nx*ny*nz cells of one expert (what TaskDataset keeps as .aabb / .cells / .cell_bounds / .cell_sizes)."""

    def __init__(self, rays: Tensor, cells: Sequence[int] = (1, 6, 6), region_bounds=None):
        dev = rays.device
        self.cells = tuple(int(c) for c in cells)
        if region_bounds is not None:                                       # :230-237 _init_region_aabb
            self.aabb = torch.tensor(region_bounds, dtype=torch.float32, device=dev)
        else:
            pts = rays[:, 0:3] + rays[:, 3:6] * rays[:, 6:7]
            self.aabb = torch.stack([pts.min(dim=0).values, pts.max(dim=0).values], dim=0)
        lo, hi = self.aabb[0], self.aabb[1]
        size = (hi - lo).clamp(min=1e-9)                                    # :174-197 _build_cell_bounds
        axes = [torch.linspace(0, 1, steps=n + 1, device=dev) for n in self.cells]
        lo_n = torch.stack(torch.meshgrid(*[a[:-1] for a in axes], indexing="ij"), dim=-1).reshape(-1, 3)
        hi_n = torch.stack(torch.meshgrid(*[a[1:] for a in axes], indexing="ij"), dim=-1).reshape(-1, 3)
        self.cell_bounds = torch.stack([lo + size * lo_n, lo + size * hi_n], dim=1).contiguous()   # [C,2,3]
        self.cell_sizes = (self.cell_bounds[:, 1] - self.cell_bounds[:, 0]).abs()
        # :241-245 _dda_transform's cell size and :595-597 the per-cell keep tolerance
        self.cell3 = torch.clamp((hi - lo) / torch.tensor(self.cells, device=dev, dtype=torch.float32), min=1e-12).contiguous()
        diag = (self.cell_bounds[:, 1] - self.cell_bounds[:, 0]).norm(dim=1)
        self.tol = torch.maximum(1e-6 * diag, torch.tensor(1e-9, device=dev)).contiguous()

    @property
    def num_cells(self) -> int:
        return self.cells[0] * self.cells[1] * self.cells[2]


def dda_route_rays(rays: Tensor, grid: TaskGrid, max_steps: int = 64, want_len: bool = False):
    """-> (cell id per ray (N,) int32, -1 = not binned; counts (C,) int32 [; in-cell length of the winning cell (N,)])."""
    rays = dev_f32(rays, "rays")
    dev, N = rays.device, rays.shape[0]
    cid = torch.empty(N, dtype=torch.int32, device=dev)
    counts = torch.zeros(grid.num_cells, dtype=torch.int32, device=dev)
    blen = torch.empty(N, dtype=torch.float32, device=dev) if want_len else None
    nx, ny, nz = grid.cells
    check(lib().acn_dda_route_rays(ctx(dev), ptr(rays), N, ptr(grid.aabb.contiguous()), nx, ny, nz, ptr(grid.cell3),
                                   ptr(grid.cell_bounds), ptr(grid.tol), int(max_steps), ptr(cid), ptr(blen), ptr(counts),
                                   stream(dev)))
    return (cid, counts, blen) if want_len else (cid, counts)


def route_and_bin(rays: Tensor, cells: Sequence[int] = (1, 6, 6), region_bounds=None, max_steps: int = 64,
                  grid: Optional[TaskGrid] = None) -> Tuple[List[Tensor], TaskGrid]:
    """TaskDataset._route_and_bin ("dda"): one int64 index tensor per cell with the rays binned there (the order inside
    a bin is unspecified -- the reference shuffles every bin right after, `_build_cell_cache`), and the grid."""
    if not rays.is_cuda:
        raise RuntimeError("route_and_bin needs CUDA rays; there is no CPU path")
    grid = grid or TaskGrid(rays, cells, region_bounds)
    cid, counts = dda_route_rays(rays, grid, max_steps)
    C = grid.num_cells
    cnt = counts.cpu()                                                      # the one host read: C ints
    offsets = torch.zeros(C, dtype=torch.int32)
    offsets[1:] = torch.cumsum(cnt, 0)[:-1].to(torch.int32)
    total = int(cnt.sum())
    sel = ops.bin_indices(cid, C, offsets.to(rays.device), total)
    off = offsets.tolist()
    return [sel[off[c]:off[c] + int(cnt[c])].long() for c in range(C)], grid
