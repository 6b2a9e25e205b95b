"""Per-kernel CUDA-event times of one routed-container training step (cfg 3 shape, one GPU)."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests" / "golden")); sys.path.insert(0, str(ROOT / "tests"))
sys.argv.append("--quick")
import synth
from adaptive_city_nerf_b200 import _lib
import importlib.util
spec = importlib.util.spec_from_file_location("bc", ROOT / "tools" / "bench_configs.py")
src = (ROOT / "tools" / "bench_configs.py").read_text().split("# ---- cfg 4")[0]
ns = {"__name__": "bc", "__file__": str(ROOT / "tools" / "bench_configs.py")}
exec(compile(src, "bench_configs_head", "exec"), ns)
container, view_rays, dev = ns["container"], ns["view_rays"], ns["dev"]
from adaptive_city_nerf_b200.nerfs.ray_rendering import render_rays
m, box = container(4, synth.CENTROIDS_G22, 1.05, False)
m.train()
N = 1 << 18
rays = torch.cat([view_rays(box, 64, 64, 60.0, seed=s)[0] for s in range(N // 4096)])
gt = torch.rand(N, 3, device=dev)
def step():
    with torch.autocast("cuda", dtype=torch.float16):
        rgb, *_ = render_rays(m, rays, ray_samples=64, active_module=None, chunk=1 << 24)
    loss = torch.nn.functional.mse_loss(rgb, gt)
    m.zero_grad(set_to_none=True)
    loss.backward()
for _ in range(2): step()
torch.cuda.synchronize()
_lib._Profile.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
prof = _lib._Profile.stop()
tot = e0.elapsed_time(e1)
print(f"step {tot:.2f} ms")
for k, (n, t) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:28s} x{n:3d}  {t:8.3f} ms  {100*t/tot:5.1f}%")
print(f"  (kernels total {sum(t for _, t in prof.values()):.2f} ms)")
