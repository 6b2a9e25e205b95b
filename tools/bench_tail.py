#!/usr/bin/env python
"""Times what surrounds the render in a training step at BASELINE configs[1] size (one expert: 64 MiB table + 14 MLP
tensors; 2^18 rays): the loss epilogue and the optimizer tail, ours vs the PyTorch calls the reference makes
(color_space_transformer + F.mse_loss; scaler.unscale_ + clip_grad_norm_ + scaler.step + scaler.update with Adam)."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "tests" / "golden")]
import synth  # noqa: E402
from adaptive_city_nerf_b200.nerfs.losses import mse_in_color_space  # noqa: E402
from adaptive_city_nerf_b200.optim import FusedAdam  # noqa: E402

dev = torch.device("cuda")
REPS = 20


def timed(fn, reps=REPS):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def torch_color_mse(pred, gt):
    gt = gt.float().clamp(0, 1)
    gt = torch.where(gt <= 0.04045, gt / 12.92, ((gt + 0.055) / 1.055).pow(2.4)).clamp(0, 1)
    return torch.nn.functional.mse_loss(pred.float().clamp(0, 1).clamp(0, 1), gt)


def main():
    out = {}
    N = 1 << 18
    pred = torch.rand(N, 3, device=dev)
    gt = torch.rand(N, 3, device=dev)

    def loss_ours():
        p = pred.clone().requires_grad_()
        mse_in_color_space(p, gt, "linear").backward()

    def loss_torch():
        p = pred.clone().requires_grad_()
        torch_color_mse(p, gt).backward()

    out["loss_fwd_bwd_ms"] = {"ours": timed(loss_ours), "torch": timed(loss_torch), "rays": N}

    shapes = [(16 << 19, 2)] + synth.expert_shapes()
    nbytes = sum(4 * torch.Size(s).numel() for s in shapes)
    mk = lambda: [torch.nn.Parameter(torch.rand(s, device=dev) - 0.5) for s in shapes]
    grads = [torch.randn(s, device=dev) * 1e-3 for s in shapes]
    for label, amp in (("fp32", False), ("amp", True)):
        a, b, c = mk(), mk(), mk()
        for ps in (a, b, c):
            for p, g in zip(ps, grads):
                p.grad = g.clone()
        ref = torch.optim.Adam(a, lr=1e-2, eps=1e-15, fused=True)
        ref_loop = torch.optim.Adam(c, lr=1e-2, eps=1e-15, foreach=True)
        ours = FusedAdam(b, lr=1e-2, eps=1e-15)
        s1, s2, s3 = (torch.amp.GradScaler("cuda", enabled=amp) for _ in range(3))
        one = torch.ones((), device=dev)
        for sc in (s1, s2, s3):
            sc.scale(one)                                   # initialise the scale tensor

        def tail_torch(opt=ref, ps=a, sc=s1):
            for p, g in zip(ps, grads):                     # stands for backward() writing fresh gradients
                p.grad.copy_(g)
            sc.unscale_(opt)
            torch.nn.utils.clip_grad_norm_(ps, 1.0)
            sc.step(opt)
            sc.update()

        def tail_ours():
            for p, g in zip(b, grads):
                p.grad.copy_(g)
            ours.step_scaled(s2, max_norm=1.0)

        def copy_only():
            for p, g in zip(b, grads):
                p.grad.copy_(g)

        base = timed(copy_only)
        t_fused = timed(tail_torch) - base
        t_foreach = timed(lambda: tail_torch(ref_loop, c, s3)) - base
        t_ours = timed(tail_ours) - base
        out[f"optimizer_tail_{label}_ms"] = {
            "ours": t_ours, "torch_fused_adam": t_fused, "torch_foreach_adam": t_foreach, "param_bytes": nbytes,
            "ours_GBps": 8 * nbytes / (t_ours * 1e-3) / 1e9,  # r g | r p,g,m,v  w p,m,v
            "what": "unscale + clip_grad_norm_(1.0) + Adam step + scaler.update, gradient copy-in subtracted"}
    out["task_binning_ms"] = bench_binning()
    print(json.dumps(out))


def torch_dda_loop(rays, t0, t1, grid, max_steps=64):
    """The reference's _dda_maxoverlap as the sequence of whole-tensor PyTorch ops it issues (restated; lower bound of
    its GPU time: region clipping, the keep test, argsort and the per-cell Python loop are not included)."""
    lo = grid.aabb[0]
    g_o, g_d = (rays[:, :3] - lo) / grid.cell3, rays[:, 3:6] / grid.cell3
    p = g_o + g_d * (t0 + 1e-6).unsqueeze(-1)
    idx = torch.floor(p).to(torch.int64)
    step = torch.sign(g_d).to(torch.int64)
    nb = torch.where(step > 0, torch.floor(p) + 1.0, torch.ceil(p) - 1.0)
    inv = 1.0 / g_d
    tmax = ((nb - p) * inv).nan_to_num_(nan=1e30, posinf=1e30, neginf=1e30)
    tdel = (step.to(p.dtype) * inv).nan_to_num_(nan=1e30, posinf=1e30, neginf=1e30)
    n = torch.tensor(grid.cells, device=rays.device)
    idx = torch.minimum(idx.clamp_min(0), n - 1)
    mul = torch.tensor([grid.cells[1] * grid.cells[2], grid.cells[2], 1], device=rays.device)
    t = t0.clone()
    best_len = torch.zeros_like(t0)
    best_cid = (idx * mul).sum(1)
    for _ in range(max_steps):
        m = tmax.min(dim=1).values
        t_next = torch.minimum(m, t1)
        dt = (t_next - t).clamp_min(0.0)
        cid = (idx * mul).sum(1)
        improve = dt > best_len
        best_len = torch.where(improve, dt, best_len)
        best_cid = torch.where(improve, cid, best_cid)
        if (t_next >= t1).all():
            break
        ax = tmax.argmin(dim=1, keepdim=True)
        adv = torch.zeros_like(idx, dtype=torch.bool).scatter_(1, ax, True)
        idx = torch.where(adv, torch.minimum((idx + step).clamp_min(0), n - 1), idx)
        tmax = torch.where(adv, tmax + tdel, tmax)
        t = t_next
    return best_cid, best_len


def bench_binning():
    import numpy as np
    from adaptive_city_nerf_b200.data import TaskGrid, route_and_bin
    from adaptive_city_nerf_b200.data.task_binning import dda_route_rays
    base = synth.task_rays(seed=91, n_soup=60000)
    N = 1 << 22
    rays = torch.from_numpy(base[np.random.default_rng(2).integers(0, base.shape[0], N)]).to(dev)
    grid = TaskGrid(rays, (1, 12, 12), tuple(map(tuple, synth.AABB_GLOBAL.tolist())))
    t_kernel = timed(lambda: dda_route_rays(rays, grid), 10)
    t_bins = timed(lambda: route_and_bin(rays, grid=grid), 5)
    t0 = rays[:, 6].clone()
    t1 = rays[:, 7].clone()
    t_torch = timed(lambda: torch_dda_loop(rays, t0, t1, grid), 2)
    return {"rays": N, "cells": grid.num_cells, "route_kernel": t_kernel, "route_and_bin": t_bins,
            "pytorch_op_sequence_dda_loop_only": t_torch, "route_kernel_GBps": N * 36 / (t_kernel * 1e-3) / 1e9}


if __name__ == "__main__":
    main()
