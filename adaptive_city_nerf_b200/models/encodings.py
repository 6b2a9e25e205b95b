"""Input encoders with the constructor / forward surface of the reference's models/encodings.py.

HashGridEncoder (reference :160-381) is the roofline-critical one: its forward and table
gradient run in csrc/hashgrid.cu with the torch branch's exact conventions (every level hashed,
`x * res` grid coordinates, floor / floor+1 corners, int64 hash == uint32 wrap-around), so hash
indices are bit-exact with the reference CPU path and fp32 features follow its un-fused lerp
order.  The parameter layout `hash_table (L*T, F)` fp32 and all buffer names match the
reference's torch branch, so checkpoints load unchanged.  There is no tinycudann branch here:
`implementation=` is accepted for signature compatibility and ignored.
"""
from __future__ import annotations

import math
import warnings
from typing import Literal, Optional

import torch
import torch.nn as nn

from .. import ops
from .._lib import INTERP

MAX_SH_DEGREE = 4
INTERPOLATIONS = ["Nearest", "Linear", "Smoothstep"]


def num_sh_bases(degree: int) -> int:
    assert degree <= MAX_SH_DEGREE, f"We don't support degree > {MAX_SH_DEGREE}."
    return (degree + 1) ** 2


# Real SH basis polynomials in (x, y, z) for l <= 4, as (coefficient, callable) rows; the l <= 3
# rows are what csrc/field_common.cuh::sh16_poly evaluates on the device (reference :27-81).
def components_from_spherical_harmonics(degree: int, directions: torch.Tensor) -> torch.Tensor:
    """Generic-degree evaluation with torch elementwise ops (API completeness; the degree-3
    case used by every config goes through the CUDA kernel, see SHEncoder.forward)."""
    assert 0 <= degree <= MAX_SH_DEGREE and directions.shape[-1] == 3
    x, y, z = directions.unbind(-1)
    xx, yy, zz = x * x, y * y, z * z
    rows = [torch.full_like(x, 0.28209479177387814)]
    if degree > 0:
        rows += [0.4886025119029199 * y, 0.4886025119029199 * z, 0.4886025119029199 * x]
    if degree > 1:
        rows += [1.0925484305920792 * x * y, 1.0925484305920792 * y * z,
                 0.9461746957575601 * zz - 0.31539156525251999, 1.0925484305920792 * x * z,
                 0.5462742152960396 * (xx - yy)]
    if degree > 2:
        rows += [0.5900435899266435 * y * (3 * xx - yy), 2.890611442640554 * x * y * z,
                 0.4570457994644658 * y * (5 * zz - 1), 0.3731763325901154 * z * (5 * zz - 3),
                 0.4570457994644658 * x * (5 * zz - 1), 1.445305721320277 * z * (xx - yy),
                 0.5900435899266435 * x * (xx - 3 * yy)]
    if degree > 3:
        rows += [2.5033429417967046 * x * y * (xx - yy), 1.7701307697799304 * y * z * (3 * xx - yy),
                 0.9461746957575601 * x * y * (7 * zz - 1), 0.6690465435572892 * y * z * (7 * zz - 3),
                 0.10578554691520431 * (35 * zz * zz - 30 * zz + 3), 0.6690465435572892 * x * z * (7 * zz - 3),
                 0.47308734787878004 * (xx - yy) * (7 * zz - 1), 1.7701307697799304 * x * z * (xx - 3 * yy),
                 0.6258357354491761 * (xx * (xx - 3 * yy) - yy * (3 * xx - yy))]
    return torch.stack(rows, dim=-1)


class SHEncoder(nn.Module):
    """Real spherical-harmonics direction encoder (reference :84-151).  levels=4 (16 components,
    the only setting any config uses) runs csrc/field_api.cu::k_sh16."""

    def __init__(self, levels: int = 4, implementation: Literal["tcnn", "torch"] = "tcnn") -> None:
        super().__init__()
        if levels <= 0 or levels > MAX_SH_DEGREE + 1:
            raise ValueError(f"Supported levels ∈ [1, {MAX_SH_DEGREE + 1}], got {levels}")
        self.levels = int(levels)
        self.degree = self.levels - 1
        self._out_dim = self.levels ** 2
        self._use_tcnn = False

    @property
    def out_dim(self) -> int:
        return self._out_dim

    def forward(self, d: torch.Tensor) -> torch.Tensor:
        assert d.shape[-1] == 3, f"Expected (...,3); got {tuple(d.shape)}"
        if self.levels == 4:
            return ops.sh16(d).to(dtype=d.dtype)      # normalises in-kernel
        dn = d / d.norm(dim=-1, keepdim=True).clamp_min(1e-9)
        return components_from_spherical_harmonics(self.degree, dn.float()).to(dtype=d.dtype)


class HashGridEncoder(nn.Module):
    """Instant-NGP multiresolution hash grid, inputs in [0,1]^3 (reference :160-381).

    forward(x (...,3)) -> (..., levels * features_per_level), differentiable w.r.t. `hash_table`.
    """

    def __init__(
        self,
        levels: int = 16,
        min_res: int = 16,
        max_res: int = 4096,
        log2_hashmap_size: int = 19,
        features_per_level: int = 2,
        hash_init_scale: float = 1e-3,
        implementation: Literal["tcnn", "torch"] = "tcnn",
        interpolation: Optional[Literal["Nearest", "Linear", "Smoothstep"]] = None,
    ) -> None:
        super().__init__()
        self.levels = int(levels)
        self.min_res = int(min_res)
        self.max_res = int(max_res)
        self.features_per_level = int(features_per_level)
        self.log2_hashmap_size = int(log2_hashmap_size)
        self.hash_init_scale = float(hash_init_scale)
        self.hash_table_size = 2 ** self.log2_hashmap_size
        self.interpolation = interpolation

        L = self.levels
        self.growth_factor = 1.0 if L <= 1 else float(
            math.exp((math.log(self.max_res) - math.log(self.min_res)) / (L - 1)))
        # fp32 evaluation on purpose: floor(16 * g^15) is 4095 in fp32, 4096 in double (reference :211-214)
        lv = torch.arange(L, dtype=torch.float32)
        res = torch.floor(self.min_res * (self.growth_factor ** lv)).to(torch.int32)
        self.register_buffer("level_resolutions", res, persistent=False)
        self._res_host = [int(r) for r in res.tolist()]      # host copy made while the buffer is still on the CPU
        self.register_buffer("level_offsets", torch.arange(L, dtype=torch.int64) * self.hash_table_size,
                             persistent=False)
        self.register_buffer("hash_primes", torch.tensor([1, 2654435761, 805459861], dtype=torch.int64),
                             persistent=False)
        self._out_dim = L * self.features_per_level
        self._use_tcnn = False

        table = (torch.rand(self.hash_table_size * L, self.features_per_level) * 2 - 1) * self.hash_init_scale
        self.hash_table = nn.Parameter(table)

        if self.interpolation is not None and self.interpolation not in INTERPOLATIONS:
            warnings.warn(f"[HashGridEncoder] interpolation '{self.interpolation}' not supported; using 'Linear'.",
                          RuntimeWarning)
            self.interpolation = "Linear"
        self._spec = None

    @property
    def out_dim(self) -> int:
        return self._out_dim

    def get_out_dim(self) -> int:
        return self._out_dim

    def grid_spec(self) -> ops.GridSpec:
        """Kernel-side description (cached; the resolution table follows the module's device)."""
        mode = INTERP[self.interpolation or "Linear"]
        s = self._spec
        if s is None or s.interp != mode or s.res.device != self.level_resolutions.device:
            s = ops.GridSpec(self.levels, self.features_per_level, self.log2_hashmap_size,
                             self.level_resolutions.to(torch.int32).contiguous(), mode, res_host=self._res_host)
            self._spec = s
        return s

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        assert x.shape[-1] == 3, f"Expected (...,3), got {tuple(x.shape)}"
        return ops.HashEncodeFn.apply(x, self.hash_table, self.grid_spec(), None).to(dtype=x.dtype)

    @torch.no_grad()
    def hash_indices(self, x: torch.Tensor) -> torch.Tensor:
        """(..., L, 8) int32 absolute table rows touched by x (corner order 000..111 = x,y,z ceil
        bits) -- the quantity the parity tests compare bit for bit."""
        x2 = x.reshape(-1, 3)
        _, idx = ops.hashgrid_fwd(x2, self.hash_table, self.grid_spec(), None, want_idx=True)
        return idx.view(*x.shape[:-1], self.levels, 8)


class FrequencyEncoder(nn.Module):
    """NeRF Fourier features (reference :387-444).  Not on any BASELINE configuration's path
    (all use the spherical direction encoding), so it stays as torch elementwise ops."""

    def __init__(self, in_dim: int, pe_dim: int, include_input: bool = True, use_pi: bool = False):
        super().__init__()
        self.in_dim = int(in_dim)
        self.pe_dim = int(pe_dim)
        self.include_input = bool(include_input)
        self.use_pi = bool(use_pi)
        self._use_tcnn = False
        self.register_buffer("bands", 2.0 ** torch.arange(self.pe_dim, dtype=torch.float32), persistent=False)

    @property
    def out_dim(self) -> int:
        return self.in_dim * (2 * self.pe_dim + (1 if self.include_input else 0))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        assert x.shape[-1] == self.in_dim, f"Expected (...,{self.in_dim}), got {tuple(x.shape)}"
        return self.torch_forward(x)

    def torch_forward(self, x: torch.Tensor) -> torch.Tensor:
        fb = self.bands.to(dtype=x.dtype, device=x.device)
        xin = x * (math.pi if self.use_pi else 1.0)
        ang = xin[..., None] * fb
        pe = torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1).reshape(*x.shape[:-1], -1)
        return torch.cat([x, pe], dim=-1) if self.include_input else pe
