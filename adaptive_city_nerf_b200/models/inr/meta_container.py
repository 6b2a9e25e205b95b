"""K spatial experts + Voronoi routing + background head (reference:
models/inr/meta_container.py:21-503).

Routing (`_routing`, stage 5) runs in csrc/routing.cu with torch.cdist's exact rounding order,
so the hard assignment and the soft support set are bit-exact w.r.t. the reference CPU path.
The reference dispatches with K rounds of nonzero()/index_select()/index_add_() (one host sync
per expert).  Here the render path (`forward_rays`) never reads anything back: a count pass leaves
the per-expert row counts on the device, `acn_bucket_plan` turns them into the bucket layout there,
the bucket pass writes capacity-bounded buckets, every expert's kernels take their row range as a
device pointer, and a blend kernel accumulates `w_k * y_k` in expert order.  (`forward(x)` on an
explicit point list still sizes its buckets with one 4K-byte host read.)"""
from __future__ import annotations

from typing import Dict, List, Literal, Optional, OrderedDict, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ..encodings import SHEncoder
from ..metamodule import MetaModule
from .meta_ngp import MetaNGP


def build_expert(nerf_variant: str, **nerf_kwargs) -> nn.Module:
    """Factory for one expert (reference :14-18).  Only the Instant-NGP variant exists here; the
    reference's 'vanilla' MetaNeRF is not constructible upstream either."""
    if nerf_variant != "instant":
        raise NotImplementedError(f"nerf_variant={nerf_variant!r}: only 'instant' (MetaNGP) is on the B200 hot path")
    return MetaNGP(**nerf_kwargs)


class MetaContainer(MetaModule):
    def __init__(
        self,
        num_submodules: int,
        centroids: torch.Tensor,
        aabb: torch.Tensor,
        nerf_variant: Literal["instant", "vanilla"] = "instant",
        boundary_margin: float = 1.0,
        cluster_2d: bool = True,
        joint_training: bool = False,
        use_bg_nerf: bool = True,
        bg_hidden: int = 32,
        bg_encoding: Literal["spherical", "fourier"] = "spherical",
        occ_conf: Optional[Dict] = None,
        **nerf_kwargs,
    ):
        super().__init__()
        assert num_submodules > 0
        assert centroids.ndim == 2 and centroids.size(0) == num_submodules
        assert boundary_margin >= 1.0
        occ_conf = occ_conf or {}
        self.register_buffer("scene_aabb_vec", torch.cat([aabb[0], aabb[1]], dim=0).float(), persistent=True)
        self.register_buffer("centroids", centroids.to(torch.float32), persistent=True)
        self.use_occ = bool(occ_conf.get("use_occ", False))
        self.boundary_margin = float(boundary_margin)
        self.cluster_2d = bool(cluster_2d)
        self.joint_training = bool(joint_training)
        self._coord_idx = (1, 2) if self.cluster_2d else (0, 1, 2)
        self.nerf_variant = nerf_variant
        self.dim_out = 4

        expert_box_list = nerf_kwargs.pop("expert_box_list")
        base = {**nerf_kwargs, "occ_conf": occ_conf}
        self.submodules = nn.ModuleList(build_expert(nerf_variant, **{**base, "scene_box": box})
                                        for box in expert_box_list)

        self.use_bg_nerf = bool(use_bg_nerf)
        if self.use_bg_nerf:
            if bg_encoding != "spherical":
                raise NotImplementedError("bg_encoding: only 'spherical' is supported")
            self.bg_dir_enc = SHEncoder(levels=4, implementation="tcnn")
            self.bg_hidden_dim = int(bg_hidden)
            # per-RAY head (N x 16 -> 32 -> 3): nn layers for the state dict; evaluated by acn_background_fwd/_bwd
            self.bg_mlp = nn.Sequential(
                nn.Linear(self.bg_dir_enc.out_dim, self.bg_hidden_dim, bias=True), nn.ReLU(),
                nn.Linear(self.bg_hidden_dim, 3, bias=True), nn.Sigmoid())

    # ------------------------------------------------------------------ routing
    def _routing(self, pts: torch.Tensor) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
        """(N,3) world points -> (weights (N,K), None) for boundary_margin > 1, else
        (None, hard (N,) int64) (reference :97-134)."""
        assert pts.dim() == 2 and pts.shape[-1] >= 3, "pts must be (N,3)"
        w, h, _ = ops.route_points(pts, self.centroids, 2 if self.cluster_2d else 3, self.boundary_margin)
        return (w, None) if w is not None else (None, h.long())

    def _sub_params(self, params):
        K = len(self.submodules)
        if params is None:
            return [None] * K
        return [self.get_subdict(params, f"submodules.{k}") for k in range(K)]

    # ------------------------------------------------------------------ field queries
    def forward(self, x: torch.Tensor, params: Optional[OrderedDict] = None,
                active_module: Optional[int] = None) -> torch.Tensor:
        """(N, D>=6) [xyz, dir, ...] -> (N,4) [rgb, sigma]; one expert if `active_module` is set,
        otherwise per-point routing and blending y = sum_k w_k y_k (reference :275-343)."""
        assert x.dim() == 2 and x.shape[-1] >= 6, "x must be (N,D>=6)"
        sub_params = self._sub_params(params)
        if active_module is not None:
            return self.submodules[active_module](x, params=sub_params[active_module])
        return self._routed(x, sub_params)

    #: the fused route-from-rays kernels keep one weight per expert in registers
    FUSED_ROUTE_MAX_EXPERTS = 16
    #: rows the routed buckets are allocated for, as a multiple of the number of samples, when boundary_margin > 1 (a
    #: sample inside the overlap band of a 2-D grid goes to 2-4 experts; with margin 1 every sample has exactly one
    #: expert and the buckets are exact).  Rows that do not fit are dropped and `check_route_overflow()` reports it.
    route_capacity_factor = 2.0

    def _overflow_flag(self, device) -> torch.Tensor:
        f = getattr(self, "_route_overflow", None)
        if f is None or f.device != device:
            f = self._route_overflow = torch.zeros(1, dtype=torch.int32, device=device)
        return f

    def check_route_overflow(self) -> None:
        """Raises if a routed batch since the last check needed more bucket rows than `route_capacity_factor` provides
        (its overflowing rows were dropped).  Reads one device word: call it outside the steps you are timing."""
        f = getattr(self, "_route_overflow", None)
        if f is not None and int(f.item()) != 0:
            f.zero_()
            raise RuntimeError("routed buckets overflowed: raise MetaContainer.route_capacity_factor "
                               f"(now {self.route_capacity_factor}); the affected batches dropped samples")

    def forward_rays(self, rays: torch.Tensor, t: torch.Tensor, params: Optional[OrderedDict] = None,
                     ray_major=False) -> torch.Tensor:
        """rays (N,8), t (N,S) -> (N,S,4): `forward(points(rays, t))` without materialising the (N*S,6) points or the
        (N*S,K) routing weights -- routing and bucketing run straight from the rays (render path of
        nerfs/ray_rendering.py:317-323 + meta_container.py:275-343)."""
        N, S = t.shape
        K = len(self.submodules)
        if K > self.FUSED_ROUTE_MAX_EXPERTS:
            return self.forward(ops.points(rays, t), params=params).view(N, S, -1)
        dims = 2 if self.cluster_2d else 3
        # ray_major (frames: consecutive rays = adjacent pixels) orders the buckets so that a warp of the experts' gathers
        # sees one sample of 32 neighbouring pixels; otherwise a ray's samples stay together (shuffled training rays).
        # bool, or a device flag from ops.rays_coherent_flag.
        P = N * S
        cap = P if self.boundary_margin <= 1.0 else int(min(float(K), float(self.route_capacity_factor)) * P)
        with torch.no_grad():
            counts, support = ops.route_count_rays(rays, t, self.centroids, dims, self.boundary_margin, want_support=True,
                                                   ray_major=ray_major)
            seg, limit, cursor = ops.bucket_plan(counts, cap, self._overflow_flag(rays.device))     # stays on the device
            sel, xd, wsel = ops.route_bucket_rays(rays, t, self.centroids, dims, self.boundary_margin, seg, cap, support=support,
                                                  ray_major=ray_major, row_limit=limit, cursor=cursor)
        y = self._evaluate_segments(xd, seg, list(range(K)), self._sub_params(params))
        return ops.BlendRangesFn.apply(y, wsel, sel, seg, P).view(N, S, -1)

    def _evaluate_segments(self, xd: torch.Tensor, seg: torch.Tensor, expert_ids: List[int], sub_params: List,
                           y_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Experts `expert_ids` on the row ranges [seg[i], seg[i+1]) of xd (device-side ranges) -> (cap,4)."""
        from .meta_ngp import autocast_half
        subs = [self.submodules[k] for k in expert_ids]
        for sub in subs:
            sub._check_fused()
        half = subs[0]._use_half(xd.device)
        experts = [(sub.xyz_encoder.grid_spec(), sub.box6()) for sub in subs]
        nodes = [ops.grad_node_of(sub.xyz_encoder.hash_table) for sub in subs]
        flat = []
        for sub, k in zip(subs, expert_ids):
            flat.append(sub.xyz_encoder.hash_table)
            flat += sub.fused_weights(sub_params[k])
        return ops.RoutedFieldFn.apply(xd, seg, half, experts, (nodes, torch.is_grad_enabled()), y_out, *flat)

    def _routed(self, x: torch.Tensor, sub_params: List) -> torch.Tensor:
        N, K = x.shape[0], len(self.submodules)
        id6 = ops.dev_f32(x[:, :6], "points")
        with torch.no_grad():
            w, hard, counts = ops.route_points(id6, self.centroids, 2 if self.cluster_2d else 3,
                                               self.boundary_margin, want_counts=True)
            cnt = counts.cpu()                       # the one host read: K ints to size the buckets
            offsets = torch.zeros(K, dtype=torch.int32)
            offsets[1:] = torch.cumsum(cnt, 0)[:-1].to(torch.int32)
            total = int(cnt.sum())
            sel, xd, wsel = ops.bucket_points(id6, w, hard, K, offsets.to(x.device), total)
        return self._evaluate_buckets(N, cnt, offsets, sel, xd, wsel, sub_params, x.device)

    def _evaluate_buckets(self, N: int, cnt, offsets, sel, xd, wsel, sub_params: List, device) -> torch.Tensor:
        out = torch.zeros(N, self.dim_out, dtype=torch.float32, device=device)
        off = offsets.tolist()
        for k, sub in enumerate(self.submodules):
            m = int(cnt[k])
            if m == 0:
                continue
            sl = slice(off[k], off[k] + m)
            yk = sub(xd[sl], params=sub_params[k])
            out = ops.BlendFn.apply(out, yk, wsel[sl], sel[sl])
        return out

    def density(self, xyz: torch.Tensor, params: Optional[OrderedDict] = None,
                active_module: Optional[int] = None) -> torch.Tensor:
        """Routed sigma (N,) (reference :217-273).  Off the render path; evaluates the experts'
        full field and keeps the sigma channel."""
        assert xyz.dim() == 2 and xyz.shape[-1] == 3, "xyz must be (N,3)"
        x6 = torch.cat([xyz, torch.zeros_like(xyz)], dim=-1)
        return self.forward(x6, params=params, active_module=active_module)[:, 3]

    def color(self, xyz: torch.Tensor, dirs: torch.Tensor, params: Optional[OrderedDict] = None,
              active_module: Optional[int] = None) -> torch.Tensor:
        """Routed rgb (N,3) (reference :137-215)."""
        assert xyz.dim() == 2 and xyz.shape[-1] == 3 and dirs.dim() == 2 and dirs.shape[-1] == 3
        x6 = torch.cat([xyz, F.normalize(dirs.to(xyz.device), dim=-1)], dim=-1)
        return self.forward(x6, params=params, active_module=active_module)[:, :3]

    # ------------------------------------------------------------------ background
    def background_color(self, d: torch.Tensor) -> torch.Tensor:
        """Background rgb per ray direction, (N,3) or (B,N,3) (reference :347-382)."""
        if not self.use_bg_nerf:
            raise RuntimeError("background_color called but use_bg_nerf=False")
        if d.dim() not in (2, 3):
            raise ValueError(f"background_color expects (N,3) or (B,N,3), got {tuple(d.shape)}")
        lin, out = self.bg_mlp[0], self.bg_mlp[2]
        if (d.is_cuda and lin.weight.is_cuda and getattr(self.bg_dir_enc, "levels", 0) == 4 and lin.in_features == 16
                and lin.out_features <= 64 and out.out_features == 3):
            # one kernel: normalise, SH16, both layers, sigmoid (csrc/background.cu); fp16 result under autocast as the
            # reference's autocast Linear gives
            from .meta_ngp import autocast_half
            d2 = d.reshape(-1, d.shape[-1])
            rgb = ops.BackgroundFn.apply(d2, lin.weight, lin.bias, out.weight, out.bias, autocast_half(d.device))
            return rgb.view(*d.shape[:-1], 3)
        dn = F.normalize(d, dim=-1).reshape(-1, 3)
        enc = self.bg_dir_enc(dn).to(dtype=lin.weight.dtype, device=lin.weight.device)
        return self.bg_mlp(enc).view(*d.shape[:-1], 3)

    # ------------------------------------------------------------------ occupancy hooks (use_occ is always False)
    def maybe_update_expert_occupancies(self, step: int, params=None) -> None:
        return None

    def freeze_expert_occupancies(self, flag: bool) -> None:
        return None

    @property
    def occ_ready(self) -> bool:
        return False

    @property
    def cells_premarked(self) -> bool:
        return False

    # ------------------------------------------------------------------ optimizer groups
    def get_param_groups(self) -> Dict[str, Dict]:
        """{"encoding","sigma","color","background"} (reference :458-503)."""
        enc: List[nn.Parameter] = []
        sig: List[nn.Parameter] = []
        col: List[nn.Parameter] = []
        for sub in self.submodules:
            g = sub.get_param_groups()
            enc += list(g["encoding"]["params"])
            sig += list(g["sigma"]["params"])
            col += list(g["color"]["params"])
        groups: Dict[str, Dict] = {}
        if enc:
            groups["encoding"] = {"params": enc}
        if sig:
            groups["sigma"] = {"params": sig}
        if col:
            groups["color"] = {"params": col}
        if self.use_bg_nerf:
            bg = list(self.bg_dir_enc.parameters()) + list(self.bg_mlp.parameters())
            if bg:
                groups["background"] = {"params": bg}
        return groups
