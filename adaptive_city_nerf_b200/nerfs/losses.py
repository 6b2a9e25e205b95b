"""Mirror of the reference's nerfs/losses.py:9-32 `compute_mse_loss`: render, bring prediction and ground truth into
one colour space (nerfs/color_space.py:22-66) and take the MSE -- the transform, the squared error, its reduction and
d loss / d pred come out of ONE kernel (`acn_color_mse`) instead of ~10 elementwise launches."""
from __future__ import annotations

from .. import ops
from .ray_rendering import render_rays


def mse_in_color_space(pred_rgb, gt_rgb, color_space: str = "linear", reduction: str = "mean"):
    """color_space_transformer + F.mse_loss.  `pred_rgb` is the rendered LINEAR colour, `gt_rgb` sRGB in [0,1]."""
    return ops.ColorMSEFn.apply(pred_rgb, gt_rgb, color_space, reduction)


def compute_mse_loss(P, model, data, params=None, active_module=None, reduction="mean"):
    """Standard MSE loss (optionally per-sample with reduction='none'); same arguments as the reference."""
    pred_rgb, *_ = render_rays(model, data["rays"], ray_samples=P.ray_samples, params=params,
                               active_module=active_module, chunk=P.chunk_points)
    return mse_in_color_space(pred_rgb, data["rgbs"], P.color_space, reduction)
