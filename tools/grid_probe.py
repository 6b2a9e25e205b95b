"""Per-level timing of the hash-grid kernels on the bench workload (run on the GPU box)."""
import sys
import json
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests" / "golden"))
import bench  # noqa: E402
from adaptive_city_nerf_b200 import ops  # noqa: E402

dev = torch.device("cuda")
rays, gt, box = bench.gpu_workload(dev, 100)
model = bench.make_model(dev, box)
ex = model.submodules[0]
S = bench.SAMPLES
jit = torch.rand(rays.shape[0], S, device=dev)
t = ops.sample_stratified(rays, S, jit)
spec = ex.xyz_encoder.grid_spec()
table = ex.xyz_encoder.hash_table.detach()
box6 = ex.box6()
P = rays.shape[0] * S


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {"P": P}
res_all = spec.res.clone()
out["fwd_all_f16"] = timeit(lambda: ops.hashgrid_fwd_rays(rays, t, table, spec, box6, torch.float16))
out["fwd_all_f32"] = timeit(lambda: ops.hashgrid_fwd_rays(rays, t, table, spec, box6, torch.float32))
dout = torch.randn(P, 32, device=dev) * 1e-3
dtab = torch.zeros_like(table)
out["bwd_all"] = timeit(lambda: ops.hashgrid_bwd_rays(rays, t, dout, spec, box6, dtab))
# shuffled point order (no ray coherence) for contrast
perm = torch.randperm(P, device=dev)
id6 = ops.points(rays, t)
xs = id6[perm].contiguous()
out["fwd_all_f16_shuffled"] = timeit(lambda: ops.hashgrid_fwd(xs, table, spec, box6, torch.float16))
out["bwd_all_shuffled"] = timeit(lambda: ops.hashgrid_bwd(xs, dout, spec, box6, dtab))
per = []
for l in range(16):
    s1 = ops.GridSpec(1, 2, spec.log2T, res_all[l:l + 1].clone(), spec.interp)
    tab1 = table[: 1 << spec.log2T].contiguous()
    d1 = dout[:, :2].contiguous()
    dt1 = torch.zeros_like(tab1)
    f = timeit(lambda: ops.hashgrid_fwd_rays(rays, t, tab1, s1, box6, torch.float16))
    b = timeit(lambda: ops.hashgrid_bwd_rays(rays, t, d1, s1, box6, dt1))
    per.append({"level": l, "res": int(res_all[l]), "fwd_ms": round(f, 3), "bwd_ms": round(b, 3)})
    print(per[-1])
out["per_level"] = per
print(json.dumps(out, indent=1))
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "grid_probe.json").write_text(json.dumps(out, indent=1))
