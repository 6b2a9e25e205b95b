"""Diagnostics for tests/test_gpu_round2.py::test_training_psnr_matches_reference_fp16_tcgen05: trajectory of OUR autocast
path against the reference's (train_amp.npz), against our fp32 path, and the eval PSNR of each."""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests" / "golden")); sys.path.insert(0, str(ROOT / "tests"))
from test_gpu_round2 import _train_setup
from helpers import cu
from adaptive_city_nerf_b200.nerfs.ray_rendering import render_rays
from adaptive_city_nerf_b200.optim import FusedAdam
from adaptive_city_nerf_b200 import ops

g = dict(np.load(ROOT / "tests/golden/train.npz"))
ref = dict(np.load(ROOT / "tests/golden/train_amp.npz"))


def run(amp, fused_bwd=True, scaler_on=True, eval_amp=None):
    ops.FUSED_EXPERT_BWD = fused_bwd
    m, pg, all_rays, gt, steps, N, S, jit = _train_setup(g)
    opt = FusedAdam(pg, eps=1e-15)
    scaler = torch.amp.GradScaler("cuda", init_scale=65536.0, enabled=scaler_on)
    m.train()
    psnr = []
    for it in range(steps):
        idx = cu(g["batch_idx"][it]).long()
        with torch.autocast("cuda", enabled=amp, dtype=torch.float16):
            rgb, *_ = render_rays(m, all_rays[idx], ray_samples=S, active_module=0, jitter=jit[it])
            loss = torch.nn.functional.mse_loss(rgb, gt[idx])
        opt.zero_grad(set_to_none=True)
        scaler.scale(loss).backward()
        scaler.step(opt, max_norm=1.0)
        scaler.update()
        psnr.append(-10.0 * np.log10(float(loss.detach()) + 1e-24))
    m.eval()
    out = {}
    for ea in (False, True):
        with torch.no_grad(), torch.autocast("cuda", enabled=ea, dtype=torch.float16):
            rgb, *_ = render_rays(m, all_rays[:2048], ray_samples=S, active_module=0)
        out[ea] = -10.0 * np.log10(float(torch.nn.functional.mse_loss(rgb, gt[:2048])) + 1e-24)
    return np.array(psnr), out


for name, kw in (("fp16 fused-bwd", dict(amp=True)), ("fp16 two-kernel bwd", dict(amp=True, fused_bwd=False)),
                 ("fp32 + scaler", dict(amp=False)), ("fp32 no scaler", dict(amp=False, scaler_on=False))):
    p, ev = run(**kw)
    d = p - ref["psnr"]
    d32 = p - g["psnr"]
    print(f"{name:22s} vs ref-amp: max|d| {np.abs(d).max():.4f} mean(last10) {d[-10:].mean():+.4f} mean(all) {d.mean():+.4f} | "
          f"vs ref-fp32 mean(last10) {d32[-10:].mean():+.4f} | at steps 10/50/100/149: {d[10]:+.4f} {d[50]:+.4f} {d[100]:+.4f} {d[149]:+.4f} | "
          f"eval fp32 {ev[False]:.4f} amp {ev[True]:.4f}  (ref amp {float(ref['final_eval_psnr']):.4f}, ref fp32 {float(g['final_eval_psnr']):.4f})")
