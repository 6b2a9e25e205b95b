"""One spatial expert: Instant-NGP style field (reference: models/inr/meta_ngp.py:15-469).

Module tree, parameter names and the `params=` fast-weight convention match the reference so
state_dicts and MAML-style callers work unchanged:

    xyz_encoder.hash_table                                   (L*T, F) fp32
    sigma_trunk.{i}.linear.{weight,bias}, sigma_head, geo_head
    color_mlp.{i}.linear.{weight,bias}, color_mlp.{depth}.{weight,bias}

`forward` -- the call the renderer makes for every sample -- runs as two kernels: hash encode
(csrc/hashgrid.cu) and the fused trunk + heads + SH + colour MLP (csrc/field_fp32.cu, or the
tcgen05 kernel csrc/field_mma.cu under torch.autocast(float16), which is how the reference selects
its fp16 path).  The occupancy-grid renderer (nerfacc, default off, partly broken upstream --
SURVEY note A) is out of scope: `use_occ=True` raises."""
from __future__ import annotations

from typing import Dict, List, Literal, Optional, Union

import torch
from torch import Tensor

from ... import ops
from ...nerfs.scene_box import SceneBox
from ..encodings import HashGridEncoder, SHEncoder
from ..metamodule import MetaLayerBlock, MetaLinear, MetaModule, MetaSequential
from ..trunc_exp import trunc_exp


def autocast_half(device: torch.device) -> bool:
    """True when the caller runs us under torch.autocast(cuda, float16) -- the reference's switch
    between its fp32 and fp16 MLP paths (meta_core.py:38, runtime_adapt.py:249-251)."""
    return (device.type == "cuda" and torch.is_autocast_enabled("cuda")
            and torch.get_autocast_dtype("cuda") == torch.float16)


class MetaNGP(MetaModule):
    def __init__(
        self,
        *,
        occ_conf: Optional[Dict],
        scene_box: SceneBox,
        hidden: int = 64,
        sigma_depth: int = 2,
        color_hidden: int = 64,
        geo_feat_dim: int = 15,
        color_depth: int = 3,
        use_sigmoid_rgb: bool = True,
        hash_enc_conf=None,
        dir_encoding: Literal["spherical", "frequency"] = "spherical",
        **kwargs,
    ) -> None:
        super().__init__()
        self.register_buffer("aabb_extent", scene_box.extent.clone())
        self.register_buffer("enc_eps", torch.tensor(1e-6, dtype=torch.float32), persistent=False)
        hash_enc_conf = hash_enc_conf or {}
        occ_conf = occ_conf or {}
        self.use_occ = bool(occ_conf.get("use_occ", False))
        if self.use_occ:
            raise NotImplementedError("use_occ=True (nerfacc occupancy renderer) is outside the B200 hot path; "
                                      "the reference's default (off) is what is implemented")
        self.geo_feat_dim = int(geo_feat_dim)
        self.use_sigmoid_rgb = bool(use_sigmoid_rgb)
        self.scene_box = scene_box
        assert isinstance(scene_box.aabb, torch.Tensor) and scene_box.aabb.shape == (2, 3)

        self.xyz_encoder = HashGridEncoder(
            levels=hash_enc_conf.get("levels", 4),
            min_res=hash_enc_conf.get("min_res", 16),
            max_res=hash_enc_conf.get("max_res", 4096),
            log2_hashmap_size=hash_enc_conf.get("log2_hashmap_size", 19),
            features_per_level=hash_enc_conf.get("features_per_level", 2),
            interpolation=hash_enc_conf.get("interpolation", "Linear"),
        )
        if dir_encoding.lower() != "spherical":
            raise NotImplementedError("only dir_encoding='spherical' (SH degree 3) is built into the fused field kernels; "
                                      "no reference config uses 'frequency'")
        self.dir_encoder = SHEncoder(levels=4)

        last = self.xyz_encoder.out_dim
        trunk = []
        for _ in range(max(int(sigma_depth), 0)):
            trunk.append(MetaLayerBlock(last, hidden, activation="relu"))
            last = hidden
        self.sigma_trunk = MetaSequential(*trunk)
        self.sigma_head = MetaLinear(last, 1)
        with torch.no_grad():
            self.sigma_head.bias.fill_(-1.0)
        self.geo_head = MetaLinear(last, self.geo_feat_dim)
        self.sigma_act = trunc_exp

        last = self.geo_feat_dim + self.dir_encoder.out_dim
        col = []
        for _ in range(max(int(color_depth), 0)):
            col.append(MetaLayerBlock(last, color_hidden, activation="relu"))
            last = color_hidden
        col.append(MetaLinear(last, 3))
        self.color_mlp = MetaSequential(*col)
        self.rgb_act = torch.nn.Sigmoid() if self.use_sigmoid_rgb else torch.nn.Identity()

        self._sigma_depth, self._color_depth = int(sigma_depth), int(color_depth)
        self._hidden, self._color_hidden = int(hidden), int(color_hidden)
        self._box6 = None

    # ------------------------------------------------------------------ kernel-side views
    #: parameter names in the order of acn_field_weights (include/acn_b200.h)
    def fused_keys(self) -> List[str]:
        return ["sigma_trunk.0.linear.weight", "sigma_trunk.0.linear.bias",
                "sigma_trunk.1.linear.weight", "sigma_trunk.1.linear.bias",
                "sigma_head.weight", "sigma_head.bias", "geo_head.weight", "geo_head.bias",
                "color_mlp.0.linear.weight", "color_mlp.0.linear.bias",
                "color_mlp.1.linear.weight", "color_mlp.1.linear.bias",
                "color_mlp.2.weight", "color_mlp.2.bias"]

    def _check_fused(self) -> None:
        if not (self._sigma_depth == 2 and self._color_depth == 2 and self._hidden == 64 and self._color_hidden == 64
                and self.use_sigmoid_rgb and self.geo_feat_dim <= 15):
            raise NotImplementedError(
                "the fused field kernels are built for the reference's configuration (hidden=color_hidden=64, "
                f"sigma_depth=color_depth=2, sigmoid rgb); got hidden={self._hidden}, color_hidden={self._color_hidden}, "
                f"sigma_depth={self._sigma_depth}, color_depth={self._color_depth}")

    def fused_weights(self, params: Optional[Dict[str, Tensor]]) -> List[Tensor]:
        """The 14 MLP tensors, taken from `params` (expert-relative keys) where present, else own
        parameters -- the reference's fallback rule (metamodule.py:61-67) without the regex walk."""
        own = dict(self.named_parameters())
        if params is None:
            return [own[k] for k in self.fused_keys()]
        return [params[k] if k in params else own[k] for k in self.fused_keys()]

    def box6(self) -> Tensor:
        """[min xyz, extent xyz] on the module's device for the in-kernel world->unit map."""
        dev = self.aabb_extent.device
        b = self._box6
        if b is None or b.device != dev:
            b = torch.cat([self.scene_box.min.to(dev).float(), self.aabb_extent.float()]).contiguous()
            self._box6 = b
        return b

    def _apply(self, fn, *a, **k):  # keep the cached box in step with .to()/.cuda()
        self._box6 = None
        return super()._apply(fn, *a, **k)

    # ------------------------------------------------------------------ reference API
    def _world_to_unit(self, x: Tensor) -> Tensor:
        x01 = (x - self.scene_box.min.to(x.device)) / self.aabb_extent
        return x01.clamp(self.enc_eps, 1.0 - self.enc_eps)

    def _enc_xyz(self, x_world: Tensor) -> Tensor:
        return ops.HashEncodeFn.apply(x_world, self.xyz_encoder.hash_table, self.xyz_encoder.grid_spec(), self.box6())

    def _enc_dir(self, d: Tensor) -> Tensor:
        return self.dir_encoder(d)

    def color(self, d: Tensor, geo_feat: Tensor, params: Optional[Dict[str, Tensor]] = None) -> Tensor:
        """Split API (reference :171-190): SH kernel + MetaLinear layers.  The renderer never calls
        this; it goes through `forward`."""
        h = torch.cat([geo_feat, self._enc_dir(d).to(geo_feat.dtype)], dim=-1)
        h = self.color_mlp(h, params=self.get_subdict(params, "color_mlp"))
        return self.rgb_act(h)

    def density(self, x: Tensor, params: Optional[Dict[str, Tensor]] = None,
                return_feats: bool = False) -> Union[Tensor, Dict[str, Tensor]]:
        """Split API (reference :192-224): hash-encode kernel + MetaLinear layers."""
        h = self._enc_xyz(x)
        h = self.sigma_trunk(h, params=self.get_subdict(params, "sigma_trunk"))
        sigma = self.sigma_act(self.sigma_head(h, params=self.get_subdict(params, "sigma_head")))
        if not return_feats:
            return sigma
        return {"sigma": sigma, "geo_feat": self.geo_head(h, params=self.get_subdict(params, "geo_head"))}

    def _use_half(self, device: torch.device) -> bool:
        """tcgen05 kernels (fp16 operands, fp32 accumulate) when the caller runs under autocast(float16) -- or when it
        has allowed TF32 matmuls (`torch.backends.cuda.matmul.allow_tf32`, which the reference's runner switches on:
        nerf_runner.py:41-42).  TF32 and fp16 carry the same 10-bit mantissa, the accumulators are fp32 either way, and
        the backward scales its gradient tiles into fp16 range, so this is the precision the reference's own GPU path
        has with that flag; without it the strict fp32 SIMT kernels run (40x / 14x slower forward / backward).  Only for
        the encoding widths the tensor-core backward is built for (16 or 32 = L*F).

        RANGE LIMIT of the TF32 stand-in: fp16 has TF32's mantissa but a 5-bit exponent -- an encoding, weight or hidden
        activation beyond +-65504 becomes inf where TF32 (8-bit exponent) would not, and magnitudes below 6e-8 flush to
        zero.  Hash features and the weights of a 64-wide ReLU MLP are O(1) (reference init: table U(+-1e-3), Linear
        default), so neither bound is near in practice; a caller who cannot rule it out sets
        `MetaNGP.tf32_on_tensor_cores = False` (class or instance) and gets the strict fp32 kernels for non-autocast
        calls.  Under autocast(float16) the reference itself has the fp16 range."""
        if self.xyz_encoder.out_dim not in (16, 32) or device.type != "cuda":
            return False
        return autocast_half(device) or (self.tf32_on_tensor_cores and bool(torch.backends.cuda.matmul.allow_tf32))

    #: serve non-autocast calls with the tcgen05 kernels (fp16 operands) when the caller allowed TF32 matmuls; see _use_half
    tf32_on_tensor_cores = True

    def forward(self, x_d: Tensor, params=None) -> Tensor:
        """(...,>=6) [xyz, dir] -> (...,4) [rgb, sigma] (reference :226-241), fused."""
        assert x_d.shape[-1] >= 6, f"Expected (...,6) [xyz,dir], got {x_d.shape}"
        self._check_fused()
        x2 = x_d.reshape(-1, x_d.shape[-1])
        table = self.xyz_encoder.hash_table
        out = ops.ExpertFieldFn.apply(x2, None, None, table, self.xyz_encoder.grid_spec(),
                                      self.box6(), self._use_half(x2.device), False, ops.grad_ctx(table),
                                      *self.fused_weights(params))
        return out.view(*x_d.shape[:-1], 4)

    def forward_rays(self, rays: Tensor, t_vals: Tensor, params=None, ray_major=False) -> Tensor:
        """Same field evaluated at the samples o + d*t of packed rays without materialising the
        (N*S,6) point tensor (nerfs/ray_rendering.py:317-319) -> (N,S,4).  `ray_major` (forward only; bool, or a device flag
        from `ops.rays_coherent_flag`): a warp of the encoder takes one sample of 32 consecutive rays -- for frames,
        where those are adjacent pixels."""
        self._check_fused()
        table = self.xyz_encoder.hash_table
        out = ops.ExpertFieldFn.apply(None, rays, t_vals, table, self.xyz_encoder.grid_spec(),
                                      self.box6(), self._use_half(rays.device), ray_major, ops.grad_ctx(table),
                                      *self.fused_weights(params))
        return out.view(t_vals.shape[0], t_vals.shape[1], 4)

    # ------------------------------------------------------------------ optimizer groups
    def get_param_groups(self) -> Dict[str, Dict]:
        """{"encoding","sigma","color"} -> {"params": [...]} (reference :446-469)."""
        return {
            "encoding": {"params": list(self.xyz_encoder.parameters())},
            "sigma": {"params": list(self.sigma_trunk.parameters()) + list(self.sigma_head.parameters())
                      + list(self.geo_head.parameters())},
            "color": {"params": list(self.color_mlp.parameters())},
        }

    # occupancy hooks the container / trainers may call; no-ops because use_occ is always False
    occ_ready = False
    occ_premarked = False
    occ_frozen = False

    def maybe_update_occ_grid(self, step: int, params=None) -> None:
        return None
