"""Fused expert backward (acn_render_expert_bwd) against the two-kernel path (acn_field_bwd -> d_enc -> acn_hashgrid_bwd_rays)
on the bench batch: CUDA-event times, and the table-gradient agreement.
    python tools/prof_fused_bwd.py [log2_rays] [--once]     (--once: one launch of each after warm-up, for ncu)"""
import json
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests" / "golden")); sys.path.insert(0, str(ROOT / "tests"))
import bench
from adaptive_city_nerf_b200 import ops

args = [a for a in sys.argv[1:] if not a.startswith("--")]
log2n = int(args[0]) if args else 18
once = "--once" in sys.argv
dev = torch.device("cuda")
rays, gt, box = bench.gpu_workload(dev, 100)
rays = rays[: 1 << log2n].contiguous()
N, S = rays.shape[0], bench.SAMPLES
model = bench.make_model(dev, box)
ex = model.submodules[0]
spec, table, box6 = ex.xyz_encoder.grid_spec(), ex.xyz_encoder.hash_table.detach(), ex.box6()
ws = [w.detach() for w in ex.fused_weights(None)]
t = ops.sample_stratified(rays, S, torch.rand(N, S, device=dev))
enc = ops.hashgrid_fwd_rays(rays, t, table, spec, box6, torch.float16)
y = ops.field_fwd(enc, rays[:, 3:], 8, S, ws, True)
bg = torch.ones(N, 3, device=dev)
rgb, dep, w, acc = ops.composite_fwd(y, t, bg, 1.0)
g_rgb = (rgb - gt[:N]) * (2.0 / rgb.numel())
d_rs = torch.empty(N, S, 4, device=dev)
ops.check(ops.lib().acn_composite_bwd(ops.ctx(dev), ops.ptr(y), ops.ptr(t), ops.ptr(bg), N, S, 1.0, ops.ptr(g_rgb), None, None, None,
                                      ops.ptr(d_rs), None, ops.stream(dev)))
d_rs = d_rs.view(-1, 4)
need = [True] * 14


def two_kernel(dt):
    grads, d_enc = ops.field_bwd(enc, rays[:, 3:], 8, S, ws, True, d_rs, True, need)
    ops.hashgrid_bwd_rays(rays, t, d_enc, spec, box6, dt)
    return grads


def fused(dt):
    return ops.render_expert_bwd(enc, (rays, t), rays[:, 3:], 8, S, ws, d_rs, need, spec, box6, dt)


def fused_single(dt):        # the single-role kernel of the debug library (the shipped kernel's predecessor)
    return ops.debug_render_expert_bwd_single(enc, (rays, t), rays[:, 3:], 8, S, ws, d_rs, need, spec, box6, dt)


def timed(fn, reps):
    dt = torch.zeros_like(table)
    for _ in range(2):
        fn(dt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn(dt)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


if once:
    d1, d2 = torch.zeros_like(table), torch.zeros_like(table)
    two_kernel(d2); fused(d1)
    torch.cuda.synchronize(); torch.cuda.profiler.start()
    two_kernel(d2); fused(d1)
    torch.cuda.synchronize(); torch.cuda.profiler.stop()
    print("ok")
else:
    d1, d2 = torch.zeros_like(table), torch.zeros_like(table)
    g2 = two_kernel(d2); g1 = fused(d1)
    rel = float((d1 - d2).double().norm() / d2.double().norm())
    relw = max(float((a - b).double().norm() / (b.double().norm() + 1e-300)) for a, b in zip(g1, g2))
    out = {"rays": N, "samples": S, "two_kernel_ms": timed(two_kernel, 10), "fused_ms": timed(fused, 10),
           "fused_single_role_ms": timed(fused_single, 10), "table_grad_rel_l2": rel, "weight_grad_rel_l2_max": relw}
    print(json.dumps(out))
