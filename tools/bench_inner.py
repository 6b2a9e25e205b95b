"""The reference's meta-training inner loop (pipelines/offline_stage/meta_core.py:14-68 task_adapt): 8 SGD steps on
the expert's 14 MLP fast weights (the hash table gets no gradient), 4000 support rays x 96 samples, autocast fp16.
Eager launches vs the same 8 steps captured in ONE CUDA graph.  Run on the GPU box."""
import json
import sys
from collections import OrderedDict
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests" / "golden"))
import bench
from adaptive_city_nerf_b200.nerfs.ray_rendering import render_rays

dev = torch.device("cuda")
rays, gt, box = bench.gpu_workload(dev, 100)
model = bench.make_model(dev, box)
ex = model.submodules[0]
N, S, STEPS, LR = 4000, 96, 8, 1e-2
rays, gt = rays[:N].contiguous(), gt[:N].contiguous()
names = [n for n, _ in ex.meta_named_parameters()]
base = OrderedDict((n, p.detach().clone()) for n, p in ex.meta_named_parameters())
jit = torch.rand(N, S, device=dev)


from adaptive_city_nerf_b200.meta import GraphedTaskAdapt, task_adapt_eager


def timed(fn, n=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


model.eval()       # fixed sample positions so that the two paths can be compared value for value (train mode re-jitters)
kw = dict(active_module=0, ray_samples=S, iterations=STEPS, inner_lr=LR)
eager_ms = timed(lambda: task_adapt_eager(model, rays, gt, **kw))
ref_fast, ref_losses = task_adapt_eager(model, rays, gt, **kw)
adapt = GraphedTaskAdapt(model, n_rays=N, **kw)
graph_ms = timed(lambda: adapt(rays, gt))
out_fast, out_losses = adapt(rays, gt)
torch.cuda.synchronize()
err = max(float((out_fast[k] - ref_fast[k].detach()).abs().max()) for k in out_fast)
model.train()
adapt_t = GraphedTaskAdapt(model, n_rays=N, **kw)
graph_train_ms = timed(lambda: adapt_t(rays, gt))
eager_train_ms = timed(lambda: task_adapt_eager(model, rays, gt, **kw))
print(json.dumps({"what": f"task_adapt: {STEPS} inner SGD steps, {N} rays x {S} samples, MLP fast weights only, autocast fp16",
                  "eval_mode": {"eager_ms_per_step": round(eager_ms / STEPS, 3), "graph_ms_per_step": round(graph_ms / STEPS, 3)},
                  "train_mode_jitter": {"eager_ms_per_step": round(eager_train_ms / STEPS, 3), "graph_ms_per_step": round(graph_train_ms / STEPS, 3)},
                  "rays_per_s_graph_train": N * STEPS / (graph_train_ms * 1e-3),
                  "max_abs_diff_adapted_weights": err, "last_loss": float(out_losses[-1])}))
if "--profile" in sys.argv:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as pr:
        adapt_t(rays, gt); torch.cuda.synchronize()
    print(pr.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
