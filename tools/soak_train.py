"""Soak: a few hundred Adam steps on an analytic scene through the tensor-core path (autocast + GradScaler, as
pipelines/online_stage/runtime_adapt.py does) -- every loss must be finite and the PSNR must climb; the same run through
the strict fp32 kernels is printed next to it.  Run on the GPU box."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests" / "golden"))
import bench
from adaptive_city_nerf_b200.nerfs.losses import mse_in_color_space
from adaptive_city_nerf_b200.nerfs.ray_rendering import render_rays
from adaptive_city_nerf_b200.optim import FusedAdam

dev = torch.device("cuda")
rays, _, box = bench.gpu_workload(dev, 7)
rays = rays[torch.randperm(rays.shape[0], device=dev, generator=torch.Generator(device=dev).manual_seed(0))][: 1 << 16].contiguous()
# analytic target: colour depends smoothly on where the ray hits the far plane
hit = rays[:, :3] + rays[:, 3:6] * rays[:, 7:8]
gt = torch.stack([0.5 + 0.5 * torch.sin(6 * hit[:, 1]), 0.5 + 0.5 * torch.cos(5 * hit[:, 2]), 0.5 + 0.25 * torch.sin(4 * hit[:, 1] * hit[:, 2])], dim=1)
STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 300
# the third run swaps PyTorch's loss / unscale / clip / Adam calls for the fused epilogue and optimizer tail ("identity":
# the analytic target is already linear, like the plain MSE of the other two runs)
for mode in ("fp16 tensor-core", "fp32 simt", "fp16 tensor-core + fused loss and optimizer tail"):
    model = bench.make_model(dev, box)
    groups = model.get_param_groups()
    fused_tail = "fused" in mode
    pg = [{"params": groups["encoding"]["params"], "lr": 1e-2}, {"params": groups["sigma"]["params"], "lr": 2e-3},
          {"params": groups["color"]["params"], "lr": 2e-3}]
    opt = FusedAdam(pg, eps=1e-15) if fused_tail else torch.optim.Adam(pg, eps=1e-15, fused=True)
    scaler = torch.amp.GradScaler("cuda", enabled=mode.startswith("fp16"))
    psnr = []
    n = STEPS if mode.startswith("fp16") else min(STEPS, 60)
    for step in range(n):
        with torch.autocast("cuda", enabled=mode.startswith("fp16"), dtype=torch.float16):
            rgb, *_ = render_rays(model, rays, ray_samples=64, active_module=0, chunk=1 << 30)
        loss = mse_in_color_space(rgb, gt, "identity") if fused_tail else torch.nn.functional.mse_loss(rgb, gt)
        opt.zero_grad(set_to_none=True)
        scaler.scale(loss).backward()
        if fused_tail:
            opt.step_scaled(scaler, max_norm=1.0)
        else:
            scaler.unscale_(opt)
            torch.nn.utils.clip_grad_norm_([p for g in opt.param_groups for p in g["params"]], 1.0)
            scaler.step(opt)
            scaler.update()
        l = float(loss.detach())
        assert np.isfinite(l), (mode, step, l)
        psnr.append(-10 * np.log10(l + 1e-24))
    marks = [0, 9, 29, 59] + ([99, 199, n - 1] if n > 60 else [])
    print(f"{mode:50s} PSNR at steps {[m + 1 for m in marks if m < n]}: " + "  ".join(f"{psnr[m]:.2f}" for m in marks if m < n) + f"   (grad scale {scaler.get_scale():.0f})")
    assert psnr[-1] > psnr[0] + 3.0, (mode, psnr[0], psnr[-1])
print("soak ok")
