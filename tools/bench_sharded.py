"""Expert-sharded rendering across GPUs (torchrun): 8 experts (2x4 grid) split over WORLD ranks, one 1920x1080 frame whose
rays are split evenly over the ranks; routed samples travel by NCCL all-to-all (adaptive_city_nerf_b200/distributed.py).
Prints one JSON line on rank 0:  torchrun --nproc-per-node W tools/bench_sharded.py [--train]"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests" / "golden")); sys.path.insert(0, str(ROOT / "tests"))
import synth
from adaptive_city_nerf_b200.distributed import ExpertShardedContainer, allreduce_grads_
from adaptive_city_nerf_b200.models.inr import MetaContainer
from adaptive_city_nerf_b200.nerfs.ray_rendering import render_rays
from adaptive_city_nerf_b200.nerfs.ray_sampling import clamp_rays_near_far, get_ray_directions, get_rays
from adaptive_city_nerf_b200.nerfs.scene_box import SceneBox

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
AABB = synth.AABB_GLOBAL
CONF = dict(levels=16, features_per_level=2, log2_hashmap_size=19, max_res=4096, min_res=16, interpolation="Linear")
ys, zs = np.linspace(-0.9, 0.9, 5)[1::2], np.linspace(-0.9, 0.9, 9)[1::2]
cen = np.array([[0.0, y, z] for y in ys for z in zs], np.float32)
torch.manual_seed(0)
box = SceneBox(T(AABB).to(dev))
full = MetaContainer(num_submodules=8, centroids=T(cen), aabb=T(AABB), boundary_margin=1.05, cluster_2d=True, use_bg_nerf=True,
                     expert_box_list=[box] * 8, hidden=64, sigma_depth=2, color_depth=2, color_hidden=64, dir_encoding="spherical",
                     hash_enc_conf=CONF, occ_conf={"use_occ": False}).to(dev)
peer = "--peer" in sys.argv
model = ExpertShardedContainer(full, peer_rows=(1 << 25) if peer else 0).shard_() if world > 1 else full
torch.cuda.empty_cache()
H, W, S = 1080, 1920, 64
cam = synth.nadir_rays(0, 1, H=H, W=W, f=1481.0 * W / 2048)[0]
dirs = get_ray_directions(H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"], True, dev)
rays = get_rays(dirs, T(cam["c2w"]).to(dev), scene_box=box).view(-1, 8)
rays, _ = clamp_rays_near_far(rays, (None, None))
n = rays.shape[0] // world
mine = rays[rank * n:(rank + 1) * n].contiguous()           # this rank's pixel rows
train = "--train" in sys.argv
if train:   # training rays come from 64 views spread over the whole scene (SURVEY 8d cfg 3), so every expert gets work
    import bench
    allr, _, _ = bench.gpu_workload(dev, 100)
    perm = torch.randperm(allr.shape[0], device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    allr = allr[perm]
    m = allr.shape[0] // world
    train_rays = allr[rank * m:(rank + 1) * m].contiguous()


def frame():
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        return render_rays(model, mine, ray_samples=S, active_module=None, chunk=1 << 24)


def step():
    sub = train_rays
    with torch.autocast("cuda", dtype=torch.float16):
        rgb, *_ = render_rays(model, sub, ray_samples=S, active_module=None, chunk=1 << 24)
    loss = rgb.square().mean()
    model.zero_grad(set_to_none=True)
    loss.backward()
    if world > 1:
        allreduce_grads_(model.shared_parameters(), average=True)


fn = step if train else frame
model.train() if train else model.eval()
for _ in range(3):
    fn()
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 6
e0.record()
for _ in range(K):
    fn()
e1.record()
dist.barrier(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / K], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if "--profile" in sys.argv:
    from adaptive_city_nerf_b200 import _lib
    from torch.profiler import profile, ProfilerActivity
    _lib._Profile.start()
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    prof = _lib._Profile.stop()
    if rank == 0:
        tot = e0.elapsed_time(e1)
        print(f"# profiled step {tot:.2f} ms; acn kernels {sum(t for _, t in prof.values()):.2f} ms", flush=True)
        for k, (c, t) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]:
            print(f"#   {k:26s} x{c:3d} {t:8.3f} ms", flush=True)
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as pr:
        fn(); torch.cuda.synchronize()
    if rank == 0:
        print("# " + pr.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=48).replace("\n", "\n# "), flush=True)
if rank == 0:
    what = (f"routed training step, 2^18 rays x {S} total" if train else f"one {W}x{H} frame, S={S}, eval fp16")
    total = (1 << 18) if train else rays.shape[0]
    print(json.dumps({"config": f"8 experts (2x4 grid, margin 1.05, bg head) sharded over {world} B200, {what}", "n_gpus": world,
                      "exchange": ("peer memory (fused dispatch / combine kernels)" if peer and world > 1 else "nccl all-to-all") if world > 1 else "none",
                      "ms": round(float(ms), 2), "samples_per_s": total * S / (float(ms) * 1e-3)}))
dist.destroy_process_group()
