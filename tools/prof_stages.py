"""One launch of every stage kernel on the bench workload shape (run under ncu on the GPU box).
    python tools/prof_stages.py [log2_rays]     default 2^18 rays x 64 samples = the bench batch"""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests" / "golden")); sys.path.insert(0, str(ROOT / "tests"))
import bench
import synth
from adaptive_city_nerf_b200 import ops

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 18
dev = torch.device("cuda")
rays, gt, box = bench.gpu_workload(dev, 100)
rays = rays[: 1 << log2n].contiguous()
N, S = rays.shape[0], bench.SAMPLES
model = bench.make_model(dev, box)
ex = model.submodules[0]
spec, table, box6 = ex.xyz_encoder.grid_spec(), ex.xyz_encoder.hash_table.detach(), ex.box6()
ws = [w.detach() for w in ex.fused_weights(None)]
jit = torch.rand(N, S, device=dev)
cen = torch.from_numpy(synth.CENTROIDS_G22).to(dev)

for rep in range(2):                       # rep 0 warms up; ncu --profile-from-start off captures rep 1 only
    if rep == 1:
        torch.cuda.synchronize(); torch.cuda.profiler.start()
    t = ops.sample_stratified(rays, S, jit)                                            # stage 1
    enc = ops.hashgrid_fwd_rays(rays, t, table, spec, box6, torch.float16)             # stage 2 fwd
    y = ops.field_fwd(enc, rays[:, 3:], 8, S, ws, True)                                # stage 3 fwd (tcgen05)
    rgb, dep, w, acc = ops.composite_fwd(y, t, torch.ones(N, 3, device=dev), 1.0)      # stage 4 fwd
    g_rgb = (rgb - gt[:N]) * (2.0 / rgb.numel())
    d_rs = torch.empty(N, S, 4, device=dev)
    ops.check(ops.lib().acn_composite_bwd(ops.ctx(dev), ops.ptr(y), ops.ptr(t), ops.ptr(torch.ones(N, 3, device=dev)), N, S, 1.0,
                                          ops.ptr(g_rgb), None, None, None, ops.ptr(d_rs), None, ops.stream(dev)))   # stage 4 bwd
    grads, d_enc = ops.field_bwd(enc, rays[:, 3:], 8, S, ws, True, d_rs.view(-1, 4), True, [True] * 14)               # stage 3 bwd
    dtable = torch.zeros_like(table)
    ops.hashgrid_bwd_rays(rays, t, d_enc, spec, box6, dtable)                          # stage 2 bwd
    id6 = ops.points(rays[: N // 8], t[: N // 8])                                      # stage 5 on an eighth of the batch
    wts, hard, counts = ops.route_points(id6, cen, 2, 1.05, want_counts=True)
    cnt = counts.cpu()
    off = torch.zeros(4, dtype=torch.int32); off[1:] = torch.cumsum(cnt, 0)[:-1].to(torch.int32)
    sel, xd, wsel = ops.bucket_points(id6, wts, hard, 4, off.to(dev), int(cnt.sum()))
    mask = ops.route_rays_voronoi(rays, 256, cen, 2, 1.1)
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", N, S)
