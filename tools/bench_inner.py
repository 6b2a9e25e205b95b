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


def adapt(fast):
    """8 first-order inner steps; returns the adapted weights and the last loss."""
    loss = None
    for _ in range(STEPS):
        with torch.autocast("cuda", dtype=torch.float16):
            rgb, *_ = render_rays(model, rays, ray_samples=S, params=fast,
                                  active_module=0, jitter=jit)
        loss = torch.nn.functional.mse_loss(rgb, gt)
        grads = torch.autograd.grad(loss, list(fast.values()))
        fast = OrderedDict((k, w - LR * g) for (k, w), g in zip(fast.items(), grads))
    return fast, loss


def fresh():
    return OrderedDict((k, v.clone().requires_grad_(True)) for k, v in base.items())


def timed(fn, n=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


model.train()
eager_ms = timed(lambda: adapt(fresh()))
ref_fast, ref_loss = adapt(fresh())

# ---- one CUDA graph for the whole adaptation
static_in = fresh()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        adapt(static_in)
torch.cuda.current_stream().wait_stream(side)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    out_fast, out_loss = adapt(static_in)


def replay():
    for k, v in base.items():
        static_in[k].data.copy_(v)
    graph.replay()


graph_ms = timed(replay)
replay(); torch.cuda.synchronize()
err = max(float((out_fast[k] - ref_fast[k]).abs().max()) for k in base)
print(json.dumps({"what": f"task_adapt: {STEPS} inner SGD steps, {N} rays x {S} samples, MLP fast weights only, autocast fp16",
                  "eager_ms": round(eager_ms, 3), "graph_ms": round(graph_ms, 3), "eager_ms_per_step": round(eager_ms / STEPS, 3),
                  "graph_ms_per_step": round(graph_ms / STEPS, 3), "rays_per_s_graph": N * STEPS / (graph_ms * 1e-3),
                  "max_abs_diff_adapted_weights": err, "loss": float(out_loss)}))
