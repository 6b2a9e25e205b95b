"""CUDA-graph capture of whole training / adaptation steps.

Every kernel of the hot path is launched without a host read -- the container's routed path included (device-side
bucket plan, device-side row ranges) -- and the optimizer tail decides skip / clip / bias corrections on the device, so
one step of the reference's online adaptation loop (pipelines/online_stage/runtime_adapt.py:287-310: compute_mse_loss ->
scaler.scale(loss).backward() -> unscale_ + clip_grad_norm_ -> scaler.step -> scaler.update) is a fixed sequence of ~150
small launches: launch-bound in eager mode (4000 rays x 96 samples over 8 experts), one graph replay here.

    step = GraphedStep(fn, example_inputs)      # fn(*tensors) -> tensor(s); warm-up, then capture
    out = step(*new_inputs)                     # copies into the static inputs, replays, returns the static outputs

Hyper-parameters passed to kernels BY VALUE (learning rates, clip norm, betas) are frozen at capture time; re-capture
after changing them (`FusedAdam` documents the same)."""
from __future__ import annotations

from typing import Callable, Sequence

import torch
from torch import Tensor


class GraphedStep:
    def __init__(self, fn: Callable, example_inputs: Sequence[Tensor], warmup: int = 3):
        self.fn = fn
        self.static_in = [t.detach().clone() for t in example_inputs]
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):             # lazy allocations (optimizer state, workspaces, caches) happen here
            for _ in range(max(1, warmup)):
                fn(*self.static_in)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = fn(*self.static_in)

    def __call__(self, *inputs: Tensor):
        for dst, src in zip(self.static_in, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out


def graphed_adapt_step(P, model, optimizer, scaler=None, grad_clip: float = 1.0, active_module=None) -> Callable:
    """The body of runtime_adapt.py:287-310 as fn(rays (N,8), rgbs (N,3)) -> loss, ready for `GraphedStep`.
    `optimizer` must be a `FusedAdam` (its step has no host read); `scaler` a torch.amp.GradScaler or None."""
    from .nerfs.losses import compute_mse_loss

    def fn(rays: Tensor, rgbs: Tensor) -> Tensor:
        optimizer.zero_grad(set_to_none=True)
        with torch.autocast("cuda", enabled=scaler is not None and scaler.is_enabled(), dtype=torch.float16):
            loss = compute_mse_loss(P, model=model, data={"rays": rays, "rgbs": rgbs}, params=None,
                                    active_module=active_module, reduction="mean")
        if scaler is not None and scaler.is_enabled():
            scaler.scale(loss).backward()
            optimizer.step_scaled(scaler, max_norm=grad_clip)
        else:
            loss.backward()
            optimizer.step(max_norm=grad_clip)
        return loss.detach()

    return fn
