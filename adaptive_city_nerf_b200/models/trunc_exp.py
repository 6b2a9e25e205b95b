"""Truncated exponential (reference: models/trunc_exp.py:32-61).

y = exp(clamp(x, -m, m)) with m = log(max finite) of the dtype; the backward differentiates the
clamped value, dy/dx = exp(clamp(x)).  On the hot path this activation is fused into the field
kernels (csrc/field_common.cuh::trunc_exp_f); this autograd version serves the split
density()/color() API."""
import torch

_LIMIT = {torch.float16: 11.089866488, torch.bfloat16: 88.722839111, torch.float32: 88.722839111,
          torch.float64: 709.782712893}


class _TruncExp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        m = _LIMIT.get(x.dtype, _LIMIT[torch.float32])
        y = torch.exp(x.clamp(-m, m))
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        return g * y


def trunc_exp(x: torch.Tensor) -> torch.Tensor:
    return _TruncExp.apply(x)
