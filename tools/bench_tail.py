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
    print(json.dumps(out))


if __name__ == "__main__":
    main()
