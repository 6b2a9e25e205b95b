"""One launch of every kernel of the round-2 training path on the bench batch (run under ncu on the GPU box): sample, fused
forward (acn_render_expert_fwd), composite fwd / bwd, loss, fused backward (acn_render_expert_bwd), optimizer tail.
    python tools/prof_stages_r2.py [log2_rays]     default 2^18 rays x 64 samples = the bench batch"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests" / "golden")); sys.path.insert(0, str(ROOT / "tests"))
import bench
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
bg = torch.ones(N, 3, device=dev)

for rep in range(2):                       # rep 0 warms up; ncu --profile-from-start off captures rep 1 only
    if rep == 1:
        torch.cuda.synchronize(); torch.cuda.profiler.start()
    t = ops.sample_stratified(rays, S, jit)                                                            # stage 1
    y, enc = ops.render_expert_fwd((rays, t), table, spec, box6, rays[:, 3:], 8, S, ws, want_enc=True)  # stages 2 + 3 forward
    rgb, dep, w, acc = ops.composite_fwd(y, t, bg, 1.0)                                                # stage 4 forward
    g_rgb = (rgb - gt[:N]) * (2.0 / rgb.numel())
    d_rs = torch.empty(N, S, 4, device=dev)
    ops.check(ops.lib().acn_composite_bwd(ops.ctx(dev), ops.ptr(y), ops.ptr(t), ops.ptr(bg), N, S, 1.0, ops.ptr(g_rgb), None, None,
                                          None, ops.ptr(d_rs), None, ops.stream(dev)))                  # stage 4 backward
    dtable = torch.zeros_like(table)
    ops.render_expert_bwd(enc, (rays, t), rays[:, 3:], 8, S, ws, d_rs.view(-1, 4), [True] * 14, spec, box6, dtable)   # stages 3 + 2 backward
    y_inf, _ = ops.render_expert_fwd((rays, t), table, spec, box6, rays[:, 3:], 8, S, ws, want_enc=False)             # inference: no encoding written
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", N, S)
