"""Inner-loop driver: first-order task adaptation replayed as ONE CUDA graph.

The reference's meta-training spends 8 of every 9 render passes in `task_adapt`
(pipelines/offline_stage/meta_core.py:14-68): `iterations` SGD steps on the expert's 14 MLP fast weights (the hash
table is not a MetaModule parameter, so it gets no gradient) over a few thousand support rays.  At that batch size the
loop is launch-bound -- ~25 kernel launches and a few hundred microseconds of Python per step -- while every kernel
behind `render_rays` is capture-safe (no host syncs; scratch is chosen at capture time).  `GraphedTaskAdapt` captures
the whole adaptation once and replays it per task: 0.99 -> 0.30 ms per inner step on a B200 (tools/bench_inner.py),
with bit-identical adapted weights.

Scope: `algo` in {fomaml, reptile} (first order, `create_graph=False`) and the MSE loss in `P.color_space`
(nerfs/losses.py:10-32 compute_mse_loss; one `acn_color_mse` launch); FIM-weighted losses and second-order MAML stay on
the eager path."""
from __future__ import annotations

from collections import OrderedDict
from typing import Callable, List, Optional, Tuple

import torch
from torch import Tensor

from .nerfs.losses import mse_in_color_space
from .nerfs.ray_rendering import render_rays


def _loss_of(loss_fn, color_space: str):
    if loss_fn is not None:
        return loss_fn
    return lambda pred, rgbs: mse_in_color_space(pred, rgbs, color_space, "mean")


def task_adapt_eager(model, rays: Tensor, rgbs: Tensor, *, active_module: int, ray_samples: int, iterations: int,
                     inner_lr: float, use_amp: bool = True, fast: Optional[OrderedDict] = None,
                     loss_fn: Optional[Callable[[Tensor, Tensor], Tensor]] = None, color_space: str = "linear",
                     chunk: int = 1 << 30, detach_updates: bool = False) -> Tuple[OrderedDict, List[Tensor]]:
    """meta_core.py:14-68 for the first-order algorithms, launch by launch (the reference's own control flow).
    The loss is compute_mse_loss's (colour-space transform + MSE, `color_space` = P.color_space) unless `loss_fn` is given.
    detach_updates=True applies each SGD step as one multi-tensor kernel outside autograd (the adapted weights then
    carry no history back to theta -- what a captured graph returns anyway)."""
    base = model.submodules[active_module]
    if fast is None:
        fast = OrderedDict((n, p) for n, p in base.meta_named_parameters())
    losses = []
    loss_fn = _loss_of(loss_fn, color_space)
    for _ in range(int(iterations)):
        with torch.autocast("cuda", enabled=use_amp, dtype=torch.float16):
            pred, *_ = render_rays(model, rays, ray_samples=ray_samples, params=fast, active_module=active_module, chunk=chunk)
            loss = loss_fn(pred, rgbs)
        grads = torch.autograd.grad(loss, tuple(fast.values()), create_graph=False, allow_unused=True)
        if detach_updates and all(g is not None for g in grads):
            with torch.no_grad():
                new = torch._foreach_add([w.detach() for w in fast.values()], [g.to(w.dtype) for g, w in zip(grads, fast.values())],
                                         alpha=-float(inner_lr))
            fast = OrderedDict((n, t.requires_grad_(True)) for n, t in zip(fast.keys(), new))
        else:
            fast = OrderedDict((n, w if g is None else (w - inner_lr * g.to(w.dtype))) for (n, w), g in zip(fast.items(), grads))
        losses.append(loss.detach())
    return fast, losses


class GraphedTaskAdapt:
    """Capture once, replay per task.

        adapt = GraphedTaskAdapt(model, active_module=cid, n_rays=4000, ray_samples=96, iterations=8, inner_lr=1e-2)
        fast, losses = adapt(support_rays, support_rgbs)      # static tensors: valid until the next call

    `fast` holds the adapted weights as plain tensors (no autograd history: the graph ran them).  For the first-order
    outer step, evaluate the query loss with `params=fast` after `fast[k].requires_grad_()` and add the resulting
    gradients to the module's parameters (`accumulate_first_order_grads`) -- in FOMAML d(fast)/d(theta) is the identity.
    """

    def __init__(self, model, *, active_module: int, n_rays: int, ray_samples: int, iterations: int, inner_lr: float,
                 use_amp: bool = True, loss_fn: Optional[Callable[[Tensor, Tensor], Tensor]] = None,
                 color_space: str = "linear", warmup: int = 3):
        self.model, self.cid = model, int(active_module)
        self.expert = model.submodules[self.cid]
        dev = next(self.expert.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedTaskAdapt needs the model on a CUDA device; there is no CPU path")
        self.names = [n for n, _ in self.expert.meta_named_parameters()]
        self.rays = torch.zeros(n_rays, 8, dtype=torch.float32, device=dev)
        self.rays[:, 5] = 1.0
        self.rays[:, 7] = 1.0
        self.rgbs = torch.zeros(n_rays, 3, dtype=torch.float32, device=dev)
        self._w = OrderedDict((n, p.detach().clone().requires_grad_(True)) for n, p in self.expert.meta_named_parameters())
        kw = dict(active_module=self.cid, ray_samples=int(ray_samples), iterations=int(iterations), inner_lr=float(inner_lr),
                  use_amp=bool(use_amp), loss_fn=loss_fn, color_space=color_space, detach_updates=True)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                       # warm-up off the capture: lazy tables, cuda attributes, allocator
            for _ in range(max(1, warmup)):
                task_adapt_eager(model, self.rays, self.rgbs, fast=self._w, **kw)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            fast, losses = task_adapt_eager(model, self.rays, self.rgbs, fast=self._w, **kw)
        self._fast = OrderedDict((n, t.detach()) for n, t in fast.items())
        self._losses = [l.detach() for l in losses]

    @torch.no_grad()
    def __call__(self, rays: Tensor, rgbs: Tensor) -> Tuple[OrderedDict, List[Tensor]]:
        if rays.shape != self.rays.shape or rgbs.shape != self.rgbs.shape:
            raise ValueError(f"captured for rays {tuple(self.rays.shape)} / rgbs {tuple(self.rgbs.shape)}")
        self.rays.copy_(rays, non_blocking=True)
        self.rgbs.copy_(rgbs, non_blocking=True)
        for (n, dst), (_, p) in zip(self._w.items(), self.expert.meta_named_parameters()):
            dst.copy_(p)                                       # theta moves between tasks (outer updates)
        self.graph.replay()
        return self._fast, self._losses


def accumulate_first_order_grads(expert, fast: OrderedDict) -> None:
    """theta.grad += d loss_query / d fast  (first-order MAML: the inner updates are treated as constants)."""
    for (n, p) in expert.meta_named_parameters():
        g = fast[n].grad
        if g is not None:
            p.grad = g.detach().clone() if p.grad is None else p.grad + g
