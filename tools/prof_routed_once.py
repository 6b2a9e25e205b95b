"""One routed training step of a 4-expert container (cfg 3 shape, 2^16 rays) and one quarter-HD frame of the 8-expert
container, after warm-up, between cudaProfilerStart/Stop -- for `ncu --profile-from-start off` captures of the routing,
bucketing, blend and per-expert range kernels.   python tools/prof_routed_once.py"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests" / "golden")); sys.path.insert(0, str(ROOT / "tests"))
src = (ROOT / "tools" / "bench_configs.py").read_text().split("# ---- cfg 4")[0]
ns = {"__name__": "bc", "__file__": str(ROOT / "tools" / "bench_configs.py")}
exec(compile(src, "bench_configs_head", "exec"), ns)
container, view_rays, grid_centroids, dev = ns["container"], ns["view_rays"], ns["grid_centroids"], ns["dev"]
import synth
from adaptive_city_nerf_b200.nerfs.ray_rendering import render_rays

m4, box = container(4, synth.CENTROIDS_G22, 1.05, False)
m4.train()
rays = torch.cat([view_rays(box, 64, 64, 60.0, seed=s)[0] for s in range(16)])
m8, _ = container(8, grid_centroids(2, 4), 1.05, True)
m8.eval()
frame_rays, _ = view_rays(box, 540, 960, 1481.0 * 960 / 2048)


def work():
    with torch.autocast("cuda", dtype=torch.float16):
        rgb, *_ = render_rays(m4, rays, ray_samples=64, active_module=None, chunk=1 << 30)
    rgb.square().mean().backward()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        render_rays(m8, frame_rays, ray_samples=64, active_module=None, chunk=1 << 30)


work(); work()
torch.cuda.synchronize(); torch.cuda.profiler.start()
work()
torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("ok")
