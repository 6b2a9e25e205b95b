"""Per-kernel CUDA-event times of one 1920x1080 routed frame (cfg 4, one GPU)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests" / "golden")); sys.path.insert(0, str(ROOT / "tests"))
from adaptive_city_nerf_b200 import _lib
src = (ROOT / "tools" / "bench_configs.py").read_text().split("# ---- cfg 4")[0]
ns = {"__name__": "bc", "__file__": str(ROOT / "tools" / "bench_configs.py")}
exec(compile(src, "bench_configs_head", "exec"), ns)
container, view_rays, grid_centroids, dev = ns["container"], ns["view_rays"], ns["grid_centroids"], ns["dev"]
from adaptive_city_nerf_b200.nerfs.ray_rendering import render_rays
m, box = container(8, grid_centroids(2, 4), 1.05, True)
m.eval()
rays, valid = view_rays(box, 1080, 1920, 1481.0 * 1920 / 2048)
def frame():
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        return render_rays(m, rays, ray_samples=64, active_module=None, chunk=1 << 24)
frame(); torch.cuda.synchronize()
_lib._Profile.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); frame(); e1.record(); torch.cuda.synchronize()
prof = _lib._Profile.stop()
tot = e0.elapsed_time(e1)
print(f"frame {tot:.2f} ms")
for k, (n, t) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:28s} x{n:3d}  {t:8.3f} ms  {100*t/tot:5.1f}%")
print(f"  (acn kernels total {sum(t for _, t in prof.values()):.2f} ms)")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as pr:
    frame(); torch.cuda.synchronize()
print(pr.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
