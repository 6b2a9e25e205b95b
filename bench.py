#!/usr/bin/env python
"""Benchmark of the B200-native rendering hot path (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
    python bench.py --impl reference [--gpus N] [--steps K] ...    # CPU arm (oracle port, host cores)

Workload (BASELINE.json configs[1]): ONE Instant-NGP expert (16-level hash grid, T = 2^19, F = 2,
64-wide MLPs), 2^18 rays x 64 samples per batch, one TRAINING step = render_rays (train mode,
stratified jitter, fp16 tcgen05 MLP under autocast) + colour-space MSE (acn_color_mse) + backward (table + 14 MLP
tensors) + global-norm clip and Adam in one pass (optim.FusedAdam).  Synthetic rays: 64 nadir 64x64 pinhole views inside the shipped scene box
(SURVEY 8d); random-init weights (reference init).  Metric: train rays/s (whole job).

N > 1 (torchrun): single-expert data parallel -- every rank renders its own 2^18-ray batch (weak
scaling) and the hash-table + MLP gradients are all-reduced over NCCL before the optimizer step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
for p in (ROOT, ROOT / "tests" / "golden"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

N_VIEWS, VIEW_HW, SAMPLES, LOG2T = 64, 64, 64, 19
N_RAYS = N_VIEWS * VIEW_HW * VIEW_HW          # 2^18
CPU_SAMPLE_RAYS = 4096                         # BASELINE configs[0] shape for the CPU arms

# algorithmic work per sample
ENC_FWD_BYTES = 16 * 8 * 2 * 4 + 64            # 1024 B gathered (fp32 table) + 64 B fp16 row written
ENC_BWD_BYTES = 2 * 1024 + 128                 # read-modify-write of the same 1024 B + 128 B fp32 dL/denc read
FIELD_FWD_FLOP = 26880
FIELD_BWD_FLOP = 2 * 26880                     # dgrad + wgrad (the recompute is not counted)
COMPOSITE_BYTES = 24
# fused backward (acn_render_expert_bwd = MLP backward + table scatter in one kernel): the 1024 B of table rows are
# read-modify-written (2048 B), the fp16 encoding row (64 B) and dL/d[rgb,sigma] (16 B) are read; d_enc never exists
FUSED_BWD_BYTES = 2 * 1024 + 64 + 16
GRID_CORNERS = 16 * 8                          # (level, corner) table rows a sample gathers forward / updates backward
# committed ncu captures (`--set full`, one launch of each kernel at this workload): kernel name in the capture per entry
# point; `traffic` = dram__bytes_read.sum + dram__bytes_write.sum of that launch is READ FROM THESE FILES
NCU_FILES = ["profiles/r02_v6_stages_ncu_full_summary.csv", "profiles/r02_v5_stages_ncu_full_summary.csv", "profiles/r02_v3_fused_bwd_ncu_full_summary.csv",
             "profiles/r01_v3_stages_ncu_full_summary.csv"]
NCU_KERNEL = {"acn_hashgrid_fwd_rays": "k_hashgrid_fwd<2, __half>", "acn_hashgrid_bwd_rays": "k_hashgrid_bwd_march<float>",
              "acn_field_fwd": "k_field_fwd_mma<32, 0>", "acn_field_bwd": "k_field_bwd_mma<32, 0>",
              "acn_render_expert_bwd": "k_expert_bwd<32>", "acn_render_expert_fwd": "k_expert_fwd<32, 16>", "acn_composite_fwd": "k_composite_fwd",
              "acn_composite_bwd": "k_composite_bwd"}


def ncu_traffic():
    """{entry point: (DRAM bytes per launch, file)} from the committed ncu summaries (first file that has the kernel)."""
    import csv
    out = {}
    for f in NCU_FILES:
        path = ROOT / f
        if not path.exists():
            continue
        rows = list(csv.reader(open(path)))
        hdr = rows[0]
        try:
            ik, ir, iw = 0, [h.startswith("dram_rd") for h in hdr].index(True), [h.startswith("dram_wr") for h in hdr].index(True)
        except ValueError:
            continue
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        ur = next((v for k, v in scale.items() if k in hdr[ir]), 1e9)
        uw = next((v for k, v in scale.items() if k in hdr[iw]), 1e9)
        for r in rows[1:]:
            for entry, kern in NCU_KERNEL.items():
                if entry not in out and r[ik] == kern:
                    out[entry] = (float(r[ir]) * ur + float(r[iw]) * uw, f)
    return out


def l2_peaks():
    """Measured ceilings of the L2-resident table accesses (tools/l2_peak.py on this pool's B200): scattered 8-byte
    gathers and red.global.add.v2/v4.f32 into a 64 MiB buffer, in accesses per second."""
    f = ROOT / "profiles" / "r02_l2_peaks.json"
    if not f.exists():
        return None
    rows = json.loads(f.read_text())["rows"]
    pick = lambda op, b: next(r["g_accesses_per_s"] for r in rows if r["buffer_mib"] == 64 and r["op"] == op and r["bytes"] == b)
    return {"gather_g_per_s": pick("gather", 8), "red_g_per_s": pick("red", 16), "source": "profiles/r02_l2_peaks.json"}


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return dict(hbm=float(d["hbm_gbs"]), tensor=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tensor=1400.0, src="fallback")


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md).  nvidia-smi needs a few hundred
    milliseconds before its first line, longer than the timed region itself, so it is started before the warm-up and only
    the lines that arrive between begin() and end() -- the timed region -- are summarised."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.t0, self.t1 = index, [], None, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc:
            self.proc.terminate()
            self.t.join(timeout=2)

    def __enter__(self):
        self.begin()
        return self

    def __exit__(self, *a):
        self.end()

    def summary(self):
        inside = [r for ts, r in self.rows if self.t0 is not None and self.t0 <= ts <= (self.t1 or ts)]
        where = "timed region"
        if not inside and self.rows and self.t0 is not None:       # region shorter than the sampling period: the line closest to it
            inside = [min(self.rows, key=lambda tr: abs(tr[0] - 0.5 * (self.t0 + (self.t1 or self.t0))))[1]]
            where = "nearest sample to the timed region (under the same load)"
        sm = [float(r[0]) for r in inside if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in inside if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in inside if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "where": where}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_workload(n_rays: int, seed: int = 0):
    import synth
    from oracle import oracle as orc
    cams = synth.nadir_rays(seed, max(1, n_rays // (VIEW_HW * VIEW_HW)), H=VIEW_HW, W=VIEW_HW, f=60.0)
    rays = []
    for cam in cams:
        d = orc.ray_directions(cam["H"], cam["W"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], True).reshape(-1, 3)
        rays.append(orc.get_rays(d, cam["c2w"], aabb=synth.AABB_GLOBAL))
    rays = np.concatenate(rays)[:n_rays]
    rays, valid = orc.clamp_near_far(rays, (None, None))
    assert valid.all()
    rng = np.random.default_rng(seed + 1)
    gt = rng.uniform(0, 1, (rays.shape[0], 3)).astype(np.float32)
    sd = synth.make_expert_params(seed + 2, log2T=LOG2T, table_scale=1e-3)
    return rays, gt, sd


def cpu_step(orc, rays, gt, sd, jitter):
    """One fwd+bwd of the hot path on the host (oracle port of the reference's torch path)."""
    import synth
    lo, hi = synth.AABB_GLOBAL
    ws = synth.expert_weight_list(sd)
    res = orc.level_resolutions()
    N = rays.shape[0]
    bg = np.ones((N, 3), np.float32)
    rgb, dep, w, acc, aux = orc.render_expert(rays, SAMPLES, ws, sd["xyz_encoder.hash_table"], lo, hi - lo, 16, 2, LOG2T,
                                              res, jitter=jitter, bg=bg)
    g_rgb = (2.0 / rgb.size) * (rgb - gt)
    d_rs, _ = orc.composite_bwd(aux["rgb_sigma"].reshape(N, SAMPLES, 4), aux["t"], bg, g_rgb)
    grads, d_enc = orc.field_bwd(aux["enc"], aux["dirs"], ws, d_rs.reshape(-1, 4))
    dt = orc.hashgrid_bwd(aux["x01"], d_enc, 16, 2, LOG2T, res)
    return float(((rgb - gt) ** 2).mean()), dt


def time_cpu(steps: int, warmup: int):
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    orc.set_threads(cores)
    orc.lib()
    rays, gt, sd = cpu_workload(CPU_SAMPLE_RAYS)
    jit = np.random.default_rng(5).uniform(0, 1, (rays.shape[0], SAMPLES)).astype(np.float32)
    for _ in range(warmup):
        cpu_step(orc, rays, gt, sd, jit)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_step(orc, rays, gt, sd, jit)
        ts.append(time.perf_counter() - t0)
    return rays.shape[0], ts, cores


def run_reference(args):
    """--impl reference: the reference algorithm's CPU implementation (oracle port: the reference is
    pure Python/torch and cannot travel to the GPU box) on all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, ts, cores = time_cpu(args.steps, max(1, min(args.warmup, 2)))
    total = sum(ts)
    val = n * len(ts) / total
    sample = f"{n} rays x {SAMPLES} samples fwd+bwd per step (configs[0] shape; T=2^{LOG2T}), no optimizer"
    print(json.dumps({
        "impl": "reference", "metric": "train rays/s", "value": val, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"single Instant-NGP expert (L16 F2 T=2^{LOG2T}, 64-wide MLPs), training step; CPU arm on a "
                               f"{n}-ray sample", "rays_per_step": n, "samples_per_ray": SAMPLES},
        "cpu_baseline": {"value": val, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------ GPU arm
def gpu_workload(dev, seed: int):
    import torch
    import synth
    from adaptive_city_nerf_b200.nerfs.ray_sampling import clamp_rays_near_far, get_ray_directions, get_rays
    from adaptive_city_nerf_b200.nerfs.scene_box import SceneBox
    box = SceneBox(torch.from_numpy(synth.AABB_GLOBAL).to(dev))
    dirs = get_ray_directions(VIEW_HW, VIEW_HW, 60.0, 60.0, VIEW_HW / 2, VIEW_HW / 2, True, dev)
    rays = []
    for cam in synth.nadir_rays(seed, N_VIEWS, H=VIEW_HW, W=VIEW_HW, f=60.0):
        rays.append(get_rays(dirs, torch.from_numpy(cam["c2w"]).to(dev), scene_box=box).view(-1, 8))
    rays = torch.cat(rays)
    rays, valid = clamp_rays_near_far(rays, (None, None))
    assert bool(valid.all()) and rays.shape[0] == N_RAYS
    gen = torch.Generator(device="cpu").manual_seed(seed + 1)
    gt = torch.rand(N_RAYS, 3, generator=gen)
    return rays, gt.to(dev), box


def make_model(dev, box):
    import torch
    import synth
    from adaptive_city_nerf_b200.models.inr import MetaContainer
    torch.manual_seed(0)
    conf = dict(levels=16, features_per_level=2, log2_hashmap_size=LOG2T, max_res=4096, min_res=16, interpolation="Linear")
    m = MetaContainer(num_submodules=1, centroids=torch.zeros(1, 3), aabb=torch.from_numpy(synth.AABB_GLOBAL),
                      boundary_margin=1.0, use_bg_nerf=False, expert_box_list=[box], hidden=64, sigma_depth=2,
                      color_depth=2, color_hidden=64, dir_encoding="spherical", hash_enc_conf=conf,
                      occ_conf={"use_occ": False}).to(dev)
    m.train()
    return m


def container_records(args, world, rank, dev):
    """BASELINE configs 3 / 4 next to the headline: (a) `expert_sharded`: a routed training step of a K-expert container
    (2x2 grid at N <= 4, 2x4 at N = 8; margin 1.05) with the experts sharded over the N ranks -- weak (2^18 rays per rank)
    and fixed-total (2^18 rays in all) -- over the peer-memory exchange, nothing read back to the host; at N = 1 the same
    container on one GPU is the baseline of the curve.  (b) `frame`: latency of one 1920x1080 frame, 8 experts with
    boundary blending and the background head, pixel rows split over the ranks with the experts replicated (a single
    view only touches the two or three experts under it, so sharding them cannot spread its work), image all-gathered."""
    import torch
    import torch.distributed as dist
    import synth
    from adaptive_city_nerf_b200.distributed import ExpertShardedContainer
    from adaptive_city_nerf_b200.models.inr import MetaContainer
    from adaptive_city_nerf_b200.nerfs.losses import mse_in_color_space
    from adaptive_city_nerf_b200.nerfs.ray_rendering import render_rays
    from adaptive_city_nerf_b200.nerfs.ray_sampling import clamp_rays_near_far, get_ray_directions, get_rays
    from adaptive_city_nerf_b200.nerfs.scene_box import SceneBox
    from adaptive_city_nerf_b200.optim import FusedAdam
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    conf = dict(levels=16, features_per_level=2, log2_hashmap_size=LOG2T, max_res=4096, min_res=16, interpolation="Linear")
    box = SceneBox(T(synth.AABB_GLOBAL).to(dev))

    def container(K, cen, use_bg):
        torch.manual_seed(0)
        return MetaContainer(num_submodules=K, centroids=T(cen), aabb=T(synth.AABB_GLOBAL), boundary_margin=1.05, cluster_2d=True,
                             use_bg_nerf=use_bg, expert_box_list=[box] * K, hidden=64, sigma_depth=2, color_depth=2, color_hidden=64,
                             dir_encoding="spherical", hash_enc_conf=conf, occ_conf={"use_occ": False}).to(dev)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n, warm=3):
        for _ in range(warm):
            fn()
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        sync()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    out = {}
    # ---------------- (a) expert-sharded routed training step
    rays_all, _, _ = gpu_workload(dev, seed=100 + rank)        # 2^18 rays of 64 views spread over the scene, this rank's own
    perm = torch.randperm(N_RAYS, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
    rays_all = rays_all[perm].contiguous()
    gt = torch.rand(N_RAYS, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(8))

    def sharded_record(K):
        cen = synth.CENTROIDS_G24 if K == 8 else synth.CENTROIDS_G22
        full = container(K, cen, False)
        if world > 1:
            model = ExpertShardedContainer(full, peer_rows=int(2.5 * N_RAYS * SAMPLES)).shard_()
            params = model.local_parameters()
        else:
            model, params = full, list(full.parameters())
        model.train()
        is_table = lambda p: p.ndim == 2 and p.shape[1] == 2 and p.shape[0] > 4096
        opt = FusedAdam([{"params": [p for p in params if is_table(p)], "lr": 1e-2},
                         {"params": [p for p in params if not is_table(p)], "lr": 2e-3}], eps=1e-15,
                        skip_zero_grads=True, norm_group=(dist.group.WORLD if world > 1 else None))
        rec = {"experts": K, "grid": "2x4" if K == 8 else "2x2", "boundary_margin": 1.05,
               "exchange": "peer memory over NVLink (kernels store rows into / load results from the owners' buffers; counts all-gathered "
                           "and laid out on the device, no host read)" if world > 1 else "none (all experts on one GPU)",
               "step": "render_rays(active_module=None) fp16 + colour-space MSE + backward + global-norm clip + Adam"}
        for tag, n in (("weak", N_RAYS), ("fixed_total", N_RAYS // world)):
            if world == 1 and tag == "fixed_total":
                rec[tag] = dict(rec["weak"])                   # the same step at N = 1
                continue
            r, g = rays_all[:n], gt[:n]

            def step():
                with torch.autocast("cuda", dtype=torch.float16):
                    rgb, *_ = render_rays(model, r, ray_samples=SAMPLES, active_module=None, chunk=1 << 30)
                loss = mse_in_color_space(rgb, g, "linear")
                opt.zero_grad(set_to_none=True)
                loss.backward()
                opt.step(max_norm=1.0)

            ms = timed(step, max(3, args.steps // 2))
            rec[tag] = {"rays_per_rank": n, "rays_total": n * world, "ms_per_step": round(ms, 3), "rays_per_s": n * world / (ms * 1e-3)}
            # what bounds a sharded step from above: the routed rows each OWNER evaluates (all ranks' rays, outside the timed region)
            with torch.no_grad():
                from adaptive_city_nerf_b200 import ops as _ops
                tt = _ops.sample_stratified(r, SAMPLES, None)
                cnt = _ops.route_count_rays(r, tt, full.centroids, 2, 1.05).to(torch.float64)
                if world > 1:
                    dist.all_reduce(cnt)
                per_owner = torch.zeros(world, dtype=torch.float64, device=dev).index_add_(0, torch.arange(K, device=dev) % world, cnt)
                rec[tag]["routed_rows_per_owner"] = {"max": int(per_owner.max()), "mean": float(per_owner.mean()),
                                                     "max_over_mean": round(float(per_owner.max() / per_owner.mean()), 3)}
            if args.graph:   # the same step as ONE CUDA graph per rank: nothing in it reads back to the host (the all-gather of
                try:         # the counts, the device-side barriers and the peer-memory kernels are captured with the rest)
                    from adaptive_city_nerf_b200.graphs import GraphedStep
                    gs = GraphedStep(lambda a, b: step(), [r, g], warmup=2)
                    msg = timed(lambda: gs(r, g), max(3, args.steps // 2), warm=2)
                    rec[tag].update(ms_per_step_cuda_graph=round(msg, 3), rays_per_s_cuda_graph=n * world / (msg * 1e-3))
                    del gs
                except Exception as e:     # noqa: BLE001 -- recorded, not fatal: the eager number stands
                    rec[tag]["cuda_graph_error"] = f"{type(e).__name__}: {str(e)[:200]}"
        sync()
        model.check_route_overflow()
        del model, full, opt, params
        torch.cuda.empty_cache()
        return rec

    out["expert_sharded"] = sharded_record(8 if world == 8 else 4)
    if world == 1:                                             # the one-GPU baseline of the 8-expert curve as well
        out["expert_sharded_8_experts"] = sharded_record(8)
    # ---------------- (b) 1080p frame, 8 experts replicated, pixel rows split over the ranks
    full = container(8, synth.CENTROIDS_G24, True).eval()
    H, W = 1080, 1920
    cam = synth.nadir_rays(0, 1, H=H, W=W, f=1481.0 * W / 2048)[0]
    dirs = get_ray_directions(H, W, cam["fx"], cam["fy"], cam["cx"], cam["cy"], True, dev)
    rays = get_rays(dirs, T(cam["c2w"]).to(dev), scene_box=box).view(-1, 8)
    rays, _ = clamp_rays_near_far(rays, (None, None))
    rows = H // world
    mine = rays[rank * rows * W:(rank + 1) * rows * W].contiguous() if rank < world - 1 else rays[rank * rows * W:].contiguous()
    img = torch.empty(world, rows * W + (H - rows * world) * W, 3, device=dev) if world > 1 else None

    def frame():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            rgb, *_ = render_rays(full, mine, ray_samples=SAMPLES, active_module=None, chunk=1 << 28)
        if world > 1:
            pad = torch.zeros(img.shape[1], 3, device=dev)
            pad[:rgb.shape[0]] = rgb
            dist.all_gather_into_tensor(img, pad)
        return rgb

    ms = timed(frame, max(3, args.steps // 2))
    sync()
    full.check_route_overflow()
    out["frame"] = {"what": "one 1920x1080 frame, 2x4 grid (8 experts, margin 1.05, background head), 64 samples per ray, eval fp16; "
                            "pixel rows split over the ranks, experts replicated, image all-gathered", "ms_per_frame": round(ms, 3),
                    "samples_per_s": H * W * SAMPLES / (ms * 1e-3)}
    del full
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from adaptive_city_nerf_b200 import _lib
    from adaptive_city_nerf_b200.nerfs.ray_rendering import render_rays
    from adaptive_city_nerf_b200.distributed import allreduce_grads_
    from adaptive_city_nerf_b200.nerfs.losses import mse_in_color_space
    from adaptive_city_nerf_b200.optim import FusedAdam

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a B200; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()

    rays, gt, box = gpu_workload(dev, seed=100 + rank)
    model = make_model(dev, box)
    groups = model.get_param_groups()
    # reference: common/utils.py get_optimizer (Adam over the encoding / sigma / color groups, configs/train.json lrs)
    opt = FusedAdam([{"params": groups["encoding"]["params"], "lr": 1e-2},
                     {"params": groups["sigma"]["params"], "lr": 2e-3},
                     {"params": groups["color"]["params"], "lr": 2e-3}], eps=1e-15)
    params = [p for g in opt.param_groups for p in g["params"]]

    def step(r, g):
        with torch.autocast("cuda", dtype=torch.float16):
            rgb, _, _, _ = render_rays(model, r, ray_samples=SAMPLES, active_module=0, chunk=1 << 30)
        loss = mse_in_color_space(rgb, g, "linear")               # nerfs/losses.py:29-32, args.py default colour space
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if world > 1:
            allreduce_grads_(params, average=True)
        opt.step(max_norm=1.0)                                    # meta_core.py:181-190 clip_all_grads + Adam, one pass
        return loss.detach()

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clk = ClockSampler(local).start()          # nvidia-smi is up and sampling by the time the timed region starts
    for _ in range(max(args.warmup, 3)):
        step(rays, gt)
    # ---- device-resident throughput (`value`): EXACTLY K steps between two events, nothing else on the stream ----
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with clk:
        e0.record()
        for _ in range(args.steps):
            step(rays, gt)
        e1.record()
        sync()
    clk.stop()
    ms = e0.elapsed_time(e1)
    # ---- the same K steps once more with a CUDA-event pair around every C-ABI call (per-kernel table, launch count);
    # the ~20 extra event records per step cost ~0.1 ms, which is why this pass is not the one `value` is taken from
    _lib._Profile.start()
    for _ in range(args.steps):
        step(rays, gt)
    sync()
    prof = _lib._Profile.stop()
    launches = _lib._Profile.launches
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    # ---- end to end through the public API with HOST buffers ----
    rays_h = rays.cpu().pin_memory()
    gt_h = gt.cpu().pin_memory()
    # Every step's inputs are copied from pinned host memory and every step's loss is read back, all inside the timed
    # region -- pipelined the way a training loop with an asynchronous loader is: the copy of step i+1's inputs runs on a
    # copy stream while step i computes, and the loss of step i is read (4 bytes into pinned memory) once step i+1 has
    # been launched, so the host never waits for the GPU with nothing queued behind it.
    copy_stream = torch.cuda.Stream(dev)
    main = torch.cuda.current_stream(dev)
    loss_h = torch.zeros(2, dtype=torch.float32).pin_memory()

    def fetch():
        with torch.cuda.stream(copy_stream):
            r, g = rays_h.to(dev, non_blocking=True), gt_h.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return r, g, ev

    def run_e2e(n):
        pending = None
        nxt = fetch()
        for i in range(n):
            r, g, ev = nxt
            main.wait_event(ev)
            r.record_stream(main); g.record_stream(main)
            nxt = fetch() if i + 1 < n else None
            loss = step(r, g)
            loss_h[i & 1].copy_(loss.detach(), non_blocking=True)          # D2H read of the step's result
            done = torch.cuda.Event()
            done.record(main)
            if pending is not None:
                pending[0].synchronize()                                    # the previous step's loss has landed
                _ = float(loss_h[pending[1]])
            pending = (done, i & 1)
        pending[0].synchronize()
        return float(loss_h[pending[1]])

    run_e2e(2)
    sync()
    t0 = time.perf_counter()
    e0.record()
    loss_val = run_e2e(args.steps)
    e1.record()
    sync()
    ms_e2e = e0.elapsed_time(e1)
    # the same loop without the pipelining (copy, step, blocking loss read, repeat), for the record
    sync()
    e0.record()
    for _ in range(args.steps):
        float(step(rays_h.to(dev, non_blocking=True), gt_h.to(dev, non_blocking=True)))
    e1.record()
    sync()
    ms_e2e_serial = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t)

    # ---- render throughput: forward only, eval mode (no jitter), no autograd ----
    model.eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        for _ in range(2):
            render_rays(model, rays, ray_samples=SAMPLES, active_module=0, chunk=1 << 30)
        sync()
        e0.record()
        for _ in range(args.steps):
            render_rays(model, rays, ray_samples=SAMPLES, active_module=0, chunk=1 << 30)
        e1.record()
        sync()
    ms_render = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_render], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_render = float(t)
    model.train()

    extra = {} if args.no_extras else container_records(args, world, rank, dev)          # expert sharding + frame latency (all ranks take part)

    if rank == 0:
        pk = peaks()
        l2 = l2_peaks()
        traffic = ncu_traffic()
        P = N_RAYS * SAMPLES
        work = {  # kernel -> (bound, algorithmic units per launch)
            "acn_hashgrid_fwd_rays": ("hbm", ENC_FWD_BYTES * P), "acn_hashgrid_bwd_rays": ("hbm", ENC_BWD_BYTES * P),
            "acn_field_fwd": ("tensor", FIELD_FWD_FLOP * P), "acn_field_bwd": ("tensor", FIELD_BWD_FLOP * P),
            "acn_render_expert_bwd": ("hbm", FUSED_BWD_BYTES * P),
            # fused forward (acn_render_expert_fwd = hash encode + MLPs): 1024 B gathered, 64 B fp16 row written for the backward
            # (training only), 16 B rgb/sigma written; the encoding is never read back
            "acn_render_expert_fwd": ("hbm", (ENC_FWD_BYTES + 16) * P),
            "acn_composite_fwd": ("hbm", COMPOSITE_BYTES * P), "acn_composite_bwd": ("hbm", (COMPOSITE_BYTES + 20) * P),
        }
        kernels = {}
        for k, (n, tot) in prof.items():
            avg = tot / max(n, 1)
            ent = {"launches": n, "avg_ms": round(avg, 4), "share": round(tot / ms, 4)}
            if k in work and avg > 0:
                bound, units = work[k]
                ach = units / (avg * 1e-3) / (1e9 if bound == "hbm" else 1e12)
                ent.update(bound=bound, achieved=round(ach, 2), frac=round(ach / pk[bound], 4),
                           unit="GB/s" if bound == "hbm" else "TFLOP/s")
            # the table is L2-resident: the honest ceilings of its accesses are the MEASURED L2 gather / atomic rates
            if l2 and avg > 0 and k in ("acn_hashgrid_fwd_rays", "acn_render_expert_fwd"):
                g = GRID_CORNERS * P / (avg * 1e-3) / 1e9
                ent["l2_gather"] = {"achieved_g_rows_per_s": round(g, 1), "peak_scattered_g_per_s": l2["gather_g_per_s"],
                                    "frac": round(g / l2["gather_g_per_s"], 3),
                                    "note": "algorithmic 8-byte corner gathers; > 1 because the coarse levels hit L1 (the peak is for scattered rows)"}
            if l2 and avg > 0 and k in ("acn_render_expert_bwd", "acn_hashgrid_bwd_rays"):
                g = GRID_CORNERS * P / (avg * 1e-3) / 1e9
                ent["l2_red"] = {"achieved_g_corner_updates_per_s": round(g, 1), "peak_g_reds_per_s": l2["red_g_per_s"],
                                 "frac": round(g / l2["red_g_per_s"], 3),
                                 "note": "algorithmic corner updates; x-neighbours share a 16-byte RED and coarse-level runs are merged, so fewer REDs are issued (ncu: 1.08 G per launch)"}
            if k == "acn_render_expert_bwd" and avg > 0:
                tf = FIELD_BWD_FLOP * P / (avg * 1e-3) / 1e12
                ent["tensor"] = {"achieved_tflops": round(tf, 1), "frac": round(tf / pk["tensor"], 4),
                                 "note": "the MLP backward inside the fused kernel (dgrad + wgrad FLOPs only)"}
            if k == "acn_render_expert_fwd" and avg > 0:
                tf = FIELD_FWD_FLOP * P / (avg * 1e-3) / 1e12
                ent["tensor"] = {"achieved_tflops": round(tf, 1), "frac": round(tf / pk["tensor"], 4),
                                 "note": "the MLP forward inside the fused kernel; the kernel is bound by the encode's gathers"}
            if k in traffic:
                ent["ncu_dram_bytes_per_launch"] = traffic[k][0]
            kernels[k] = ent
        top = max((k for k in kernels if "bound" in kernels[k]), key=lambda k: kernels[k]["avg_ms"] * kernels[k]["launches"])
        tk = kernels[top]
        roof = {"kernel": top, "bound": tk["bound"], "achieved": tk["achieved"], "peak": pk[tk["bound"]], "unit": tk["unit"],
                "frac": tk["frac"], "traffic": traffic.get(top, (None, None))[0],
                "traffic_source": f"dram__bytes_read.sum + dram__bytes_write.sum of one launch, read from {traffic[top][1]}" if top in traffic else None,
                "algorithmic": work[top][1], "peak_source": pk["src"], "avg_ms": tk["avg_ms"]}
        for alt in ("l2_red", "l2_gather", "tensor"):
            if alt in tk:
                roof[alt] = tk[alt]
        n_cpu, ts, cores = time_cpu(steps=2, warmup=1) if (world == 1 and not args.no_extras) else (0, [], 0)
        out = {
            "metric": "train rays/s", "value": world * N_RAYS * args.steps / (ms * 1e-3), "unit": "rays/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": f"single Instant-NGP expert (L16 F2 T=2^{LOG2T}, 64-wide MLPs), 2^18 rays/batch x {SAMPLES} "
                                   "samples, training step = render_rays fwd + colour-space MSE + bwd + grad clip + Adam (configs[1])",
                       "rays_per_step_per_gpu": N_RAYS, "samples_per_ray": SAMPLES, "parallelism": f"dp{world}",
                       "l2": "inputs > L2: 64 MiB table + 1 GiB fp16 encodings (written forward, read by the fused backward) + 256 MiB "
                             "rgb/sigma + 256 MiB of their gradients per step; d_enc is never materialised"},
            "samples_per_s": world * N_RAYS * SAMPLES * args.steps / (ms * 1e-3),
            "e2e": {"value": world * N_RAYS * args.steps / (ms_e2e * 1e-3), "unit": "rays/s",
                    "h2d_bytes_per_step": int(rays_h.numel() * 4 + gt_h.numel() * 4), "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps, "last_loss": loss_val,
                    "how": "every step: H2D of that step's rays + targets from pinned memory, the step through render_rays / "
                           "compute loss / FusedAdam, D2H of its loss; step i+1's copy runs on a copy stream under step i and "
                           "step i's loss is read once step i+1 is launched",
                    "ms_per_step_unpipelined": round(ms_e2e_serial / args.steps, 4)},
            "gpu_launches": launches, "clocks": clk.summary(), "roofline": roof, "kernels": kernels,
            "render": {"metric": "render samples/s", "value": world * N_RAYS * SAMPLES * args.steps / (ms_render * 1e-3),
                       "unit": "samples/s", "ms_per_batch": ms_render / args.steps,
                       "what": "render_rays forward only (eval, no_grad, autocast fp16), same expert and rays"},
        }
        out.update(extra)
        if ts:
            out["cpu_baseline"] = {"value": n_cpu / min(ts), "unit": "rays/s", "cores": cores, "kind": "port",
                                   "sample": f"{n_cpu} rays x {SAMPLES} samples fwd+bwd (configs[0] shape, T=2^{LOG2T}), "
                                             "best of 2, no optimizer"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="headline step only: skip the container / frame records and the CPU baseline (kernel A/B runs)")
    ap.add_argument("--graph", action="store_true", help="also time the expert-sharded step as one CUDA graph per rank")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
