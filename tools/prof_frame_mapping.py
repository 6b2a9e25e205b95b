"""One chunk of the 1080p 8-expert frame rendered twice -- sample-major and ray-major warp mapping -- between
cudaProfilerStart/Stop, for `ncu --set full -k regex:k_hashgrid_fwd` (tools/gpu_prof_frame.sh).  Plain run: prints the
chunk's hash-encode time per mapping."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests" / "golden")); sys.path.insert(0, str(ROOT / "tests"))
from adaptive_city_nerf_b200 import _lib
src = (ROOT / "tools" / "bench_configs.py").read_text().split("# ---- cfg 4")[0]
ns = {"__name__": "bc", "__file__": str(ROOT / "tools" / "bench_configs.py")}
exec(compile(src, "bench_configs_head", "exec"), ns)
container, view_rays, grid_centroids = ns["container"], ns["view_rays"], ns["grid_centroids"]
from adaptive_city_nerf_b200.nerfs.ray_rendering import render_rays
m, box = container(8, grid_centroids(2, 4), 1.05, True)
m.eval()
rays, _ = view_rays(box, 1080, 1920, 1481.0 * 1920 / 2048)
chunk = rays[540 * 1920: 540 * 1920 + (1 << 18)].contiguous()          # 2^18 rays from the middle rows = 2^24 samples


def go(coherent):
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        return render_rays(m, chunk, ray_samples=64, active_module=None, chunk=1 << 30, coherent_rays=coherent)


for c in (False, True):
    go(c)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for c in (False, True):
    _lib._Profile.start()
    go(c)
    torch.cuda.synchronize()
    prof = _lib._Profile.stop()
    n, t = prof["acn_hashgrid_fwd"]
    print(f"coherent_rays={c}: acn_hashgrid_fwd x{n} {t:.3f} ms; all acn kernels {sum(v[1] for v in prof.values()):.3f} ms", flush=True)
torch.cuda.profiler.stop()
