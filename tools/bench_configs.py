"""BASELINE.json configs[2..4] on ONE GPU (all experts local): routed-container training step (cfg 3 shape), 1920x1080
viewer frame latency with 8 experts (cfg 4), online-adaptation steps (cfg 5).  Prints one JSON line per config.
Run on the GPU box:  python tools/bench_configs.py [--quick]"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests" / "golden")); sys.path.insert(0, str(ROOT / "tests"))
import synth
from adaptive_city_nerf_b200.models.inr import MetaContainer
from adaptive_city_nerf_b200.nerfs.ray_rendering import render_rays
from adaptive_city_nerf_b200.nerfs.losses import mse_in_color_space
from adaptive_city_nerf_b200.optim import FusedAdam
from adaptive_city_nerf_b200.nerfs.ray_sampling import clamp_rays_near_far, get_ray_directions, get_rays
from adaptive_city_nerf_b200.nerfs.scene_box import SceneBox

quick = "--quick" in sys.argv
dev = torch.device("cuda")
T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
AABB = synth.AABB_GLOBAL
CONF = dict(levels=16, features_per_level=2, log2_hashmap_size=19, max_res=4096, min_res=16, interpolation="Linear")


def grid_centroids(ny, nz):
    """scripts/create_clusters.py:298-314 _grid_centroids over the camera extent used by the synthetic views."""
    ys = np.linspace(-0.9, 0.9, 2 * ny + 1)[1::2]
    zs = np.linspace(-0.9, 0.9, 2 * nz + 1)[1::2]
    return np.array([[0.0, y, z] for y in ys for z in zs], np.float32)


def container(K, cen, margin, use_bg):
    torch.manual_seed(0)
    box = SceneBox(T(AABB).to(dev))
    m = MetaContainer(num_submodules=K, centroids=T(cen), aabb=T(AABB), boundary_margin=margin, cluster_2d=True,
                      use_bg_nerf=use_bg, expert_box_list=[box] * K, hidden=64, sigma_depth=2, color_depth=2, color_hidden=64,
                      dir_encoding="spherical", hash_enc_conf=CONF, occ_conf={"use_occ": False}).to(dev)
    return m, box


def view_rays(box, H, W, f, seed=0):
    cam = synth.nadir_rays(seed, 1, H=H, W=W, f=f)[0]
    dirs = get_ray_directions(H, W, f, f, W / 2, H / 2, True, dev)
    rays = get_rays(dirs, T(cam["c2w"]).to(dev), scene_box=box).view(-1, 8)
    rays, valid = clamp_rays_near_far(rays, (None, None))
    return rays, valid


def timed(fn, n, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


# ---- cfg 4: viewer frame, 2x4 grid, 8 experts, boundary blending, bg head
m, box = container(8, grid_centroids(2, 4), 1.05, True)
m.eval()
H, W = (540, 960) if quick else (1080, 1920)
rays, valid = view_rays(box, H, W, 1481.0 * W / 2048)
S = 64
with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
    res4 = {}
    for chunk in (1 << 24, 1 << 28):
        frames = [timed(lambda: render_rays(m, rays, ray_samples=S, active_module=None, chunk=chunk), 1, warm=2 if i == 0 else 0)
                  for i in range(6)]
        res4[chunk] = (sorted(frames)[len(frames) // 2], frames)
    ms, frames = res4[1 << 24]
print(json.dumps({"config": "cfg4: 2x4 grid, 8 experts, margin 1.05, bg head, one %dx%d frame, S=64, eval fp16, 1 GPU" % (W, H),
                  "ms_per_frame": round(ms, 2), "samples_per_s": rays.shape[0] * S / (ms * 1e-3), "valid_rays": int(valid.sum()),
                  "frames_ms": [round(f, 2) for f in frames], "what": "median of 6 frames after 2 warm-up frames, chunk = 2^24 points",
                  "ms_per_frame_whole_frame_in_one_chunk": round(res4[1 << 28][0], 2)}))
del m
torch.cuda.empty_cache()

# ---- cfg 3 shape on one GPU: 2x2 grid, 4 experts, routed training step
m, box = container(4, synth.CENTROIDS_G22, 1.05, False)
m.train()
N = 1 << (16 if quick else 18)
rays = torch.cat([view_rays(box, 64, 64, 60.0, seed=s)[0] for s in range(N // 4096)])
gt = torch.rand(N, 3, device=dev)
groups = m.get_param_groups()
opt = FusedAdam([{"params": groups["encoding"]["params"], "lr": 1e-2}, {"params": groups["sigma"]["params"], "lr": 2e-3},
                 {"params": groups["color"]["params"], "lr": 2e-3}], eps=1e-15)


def step3():
    with torch.autocast("cuda", dtype=torch.float16):
        rgb, *_ = render_rays(m, rays, ray_samples=64, active_module=None, chunk=1 << 24)
    loss = mse_in_color_space(rgb, gt, "linear")
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step(max_norm=1.0)


ms = timed(step3, 3, warm=2)
print(json.dumps({"config": "cfg3 shape on 1 GPU: 2x2 grid, 4 experts, margin 1.05, routed training step, %d rays x 64" % N,
                  "ms_per_step": round(ms, 2), "rays_per_s": N / (ms * 1e-3)}))
del m, opt
torch.cuda.empty_cache()

# ---- cfg 5: online adaptation, 8 experts, 4000 support rays x 96 samples, Adam + GradScaler + clip
m, box = container(8, grid_centroids(2, 4), 1.05, True)
m.train()
rays = view_rays(box, 64, 64, 60.0, seed=3)[0][:4000].contiguous()
gt = torch.rand(rays.shape[0], 3, device=dev)
groups = m.get_param_groups()
pg = [{"params": groups["encoding"]["params"], "lr": 1e-2}, {"params": groups["sigma"]["params"], "lr": 2e-3},
      {"params": groups["color"]["params"], "lr": 2e-3}, {"params": groups["background"]["params"], "lr": 1e-3}]
all_params = [p for g in pg for p in g["params"]]
res5 = {}
for tail in ("torch", "fused"):
    opt = torch.optim.Adam(pg, eps=1e-15, fused=True) if tail == "torch" else FusedAdam(pg, eps=1e-15)
    scaler = torch.amp.GradScaler("cuda")

    def step5():
        with torch.autocast("cuda", dtype=torch.float16):
            rgb, *_ = render_rays(m, rays, ray_samples=96, active_module=None, chunk=1 << 24)
            loss = torch.nn.functional.mse_loss(rgb.float(), gt) if tail == "torch" else mse_in_color_space(rgb, gt, "linear")
        opt.zero_grad(set_to_none=True)
        scaler.scale(loss).backward()
        if tail == "torch":                                  # runtime_adapt.py:262-268 as written
            scaler.unscale_(opt)
            torch.nn.utils.clip_grad_norm_(all_params, 1.0)
            scaler.step(opt)
            scaler.update()
        else:
            opt.step_scaled(scaler, max_norm=1.0)

    res5[tail] = timed(step5, 8, warm=3)
    del opt
    torch.cuda.empty_cache()
# the same step as ONE CUDA graph (nothing in it reads back to the host any more)
import types
from adaptive_city_nerf_b200.graphs import GraphedStep, graphed_adapt_step
from adaptive_city_nerf_b200.optim import get_optimizer
Pn = types.SimpleNamespace(lr=1e-3, encoding_lr=1e-2, sigma_lr=2e-3, color_lr=2e-3, bg_lr=1e-3, optimizer="adam", weight_decay=0.0,
                           ray_samples=96, chunk_points=1 << 24, color_space="linear")
opt = get_optimizer(Pn, m)
scaler = torch.amp.GradScaler("cuda")
gstep = GraphedStep(graphed_adapt_step(Pn, m, opt, scaler, grad_clip=1.0), [rays, gt])
res5["graph"] = timed(lambda: gstep(rays, gt), 20, warm=3)
m.check_route_overflow()
ms = res5["fused"]
print(json.dumps({"config": "cfg5: 8 experts, 4000 support rays x 96 samples, Adam + GradScaler + clip 1.0, whole container",
                  "ms_per_step": round(ms, 2), "rays_per_s": rays.shape[0] / (ms * 1e-3),
                  "ms_per_step_with_pytorch_loss_and_optimizer": round(res5["torch"], 2),
                  "ms_per_step_cuda_graph": round(res5["graph"], 3), "rays_per_s_cuda_graph": rays.shape[0] / (res5["graph"] * 1e-3)}))
