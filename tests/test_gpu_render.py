"""End-to-end parity of render_rays against outputs of the unmodified reference (golden fixtures):
one expert (eval / train with shared jitter / fast weights), the routed 4-expert container with
the background head, and a fixed-step training run (PSNR within 0.1 dB)."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

import synth
from helpers import F32, cu, make_container, npy, rel_err

pytestmark = pytest.mark.gpu


def _render():
    from adaptive_city_nerf_b200.nerfs.ray_rendering import render_rays
    return render_rays


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_single_expert_fp32(golden, mode):
    g = golden("render")
    render_rays = _render()
    m = make_container(1, np.zeros((1, 3)), [synth.AABB_GLOBAL], 1.0, False, 83)
    m.train(mode == "train")
    jit = cu(g["jitter"]) if mode == "train" else None
    rgb, dep, w, acc = render_rays(m, cu(g["rays"]), ray_samples=32, active_module=0, jitter=jit)
    # fp32 path: per-pixel rgb/depth within 1e-5 abs (north-star bar is 1e-3)
    np.testing.assert_allclose(npy(rgb), g[f"{mode}.rgb"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(npy(dep), g[f"{mode}.depth"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(npy(w), g[f"{mode}.weights"], atol=1e-5, rtol=1e-4)
    np.testing.assert_allclose(npy(acc), g[f"{mode}.acc"], atol=1e-5, rtol=0)
    loss = (rgb * cu(g["G_rgb"])).sum() + (dep * cu(g["G_depth"])).sum()
    loss.backward()
    named = dict(m.submodules[0].named_parameters())
    for key in synth.EXPERT_KEYS:
        assert rel_err(npy(named[key].grad), g[f"{mode}.grad.{key}"]) < 2e-4, key
    tg = npy(m.submodules[0].xyz_encoder.hash_table.grad)
    assert rel_err(tg[::97], g[f"{mode}.grad.table_sub"]) < 2e-4
    dig = g[f"{mode}.grad.table_digest"]
    assert abs(np.abs(tg).sum(dtype=np.float64) - dig[1]) / dig[1] < 1e-4


def test_single_expert_fp16_autocast(golden):
    """autocast(fp16) selects the tcgen05 MLP: per-pixel rgb/depth within 1e-3 abs of the fp32 reference."""
    g = golden("render")
    render_rays = _render()
    m = make_container(1, np.zeros((1, 3)), [synth.AABB_GLOBAL], 1.0, False, 83).eval()
    with torch.autocast("cuda", dtype=torch.float16):
        rgb, dep, w, acc = render_rays(m, cu(g["rays"]), ray_samples=32, active_module=0)
    assert rgb.dtype == torch.float32
    assert np.abs(npy(rgb) - g["eval.rgb"]).max() < 1e-3
    assert np.abs(npy(dep) - g["eval.depth"]).max() < 1e-3
    loss = (rgb * cu(g["G_rgb"])).sum() + (dep * cu(g["G_depth"])).sum()
    loss.backward()
    named = dict(m.submodules[0].named_parameters())
    # gradients of the fp16 forward vs the fp32 reference: relative L2 <= 2e-2.  (The first-layer weight
    # gradient measured 1.04e-2 on B200: it sees the fp16 rounding of O(0.5) test encodings directly.)
    for key in synth.EXPERT_KEYS:
        a, b = npy(named[key].grad).ravel(), g[f"eval.grad.{key}"].ravel()
        assert np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30) < 2e-2, key


def test_fast_weights_params(golden):
    """params= (expert-relative keys when active_module is set): values and grads w.r.t. the fast weights;
    the module's own parameters must receive no gradient (FOMAML inner loop, meta_core.py:54-59)."""
    g = golden("render")
    render_rays = _render()
    m = make_container(1, np.zeros((1, 3)), [synth.AABB_GLOBAL], 1.0, False, 83).eval()
    fast = OrderedDict((n, (p * 1.25).detach().requires_grad_(True)) for n, p in m.submodules[0].meta_named_parameters())
    assert list(fast) == [str(k) for k in g["fast.keys"]]
    rgb, *_ = render_rays(m, cu(g["rays"]), ray_samples=32, params=fast, active_module=0)
    np.testing.assert_allclose(npy(rgb), g["fast.rgb"], atol=1e-5, rtol=0)
    grads = torch.autograd.grad((rgb * cu(g["G_rgb"])).sum(), list(fast.values()))
    for (n, _), gi in zip(fast.items(), grads):
        assert rel_err(npy(gi), g["fast.grad." + n]) < 2e-4, n
    assert all(p.grad is None for p in m.parameters())


@pytest.mark.parametrize("tag,margin", [("soft", 1.05), ("hard", 1.0)])
def test_container_routed(golden, tag, margin):
    g = golden("render")
    render_rays = _render()
    m = make_container(4, synth.CENTROIDS_G22, synth.EXPERT_BOXES_G22, margin, True, 85).eval()
    rays = cu(g["rays4"])
    rgb, dep, w, acc = render_rays(m, rays, ray_samples=32, active_module=None)
    np.testing.assert_allclose(npy(rgb), g[f"{tag}.rgb"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(npy(dep), g[f"{tag}.depth"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(npy(acc), g[f"{tag}.acc"], atol=1e-5, rtol=0)
    (rgb * cu(g["G4"])).sum().backward()
    for k in range(4):
        named = dict(m.submodules[k].named_parameters())
        for key in ("sigma_trunk.0.linear.weight", "color_mlp.2.bias", "sigma_head.weight"):
            assert rel_err(npy(named[key].grad), g[f"{tag}.grad.{k}.{key}"]) < 2e-4, (k, key)
        tg = npy(m.submodules[k].xyz_encoder.hash_table.grad)
        ref = g[f"{tag}.grad.{k}.table_sub"]
        assert np.abs(tg[::97] - ref).max() <= 2e-4 * np.abs(ref).max() + 1e-9, k
    for key in ("bg_mlp.0.weight", "bg_mlp.2.bias"):
        assert rel_err(npy(dict(m.named_parameters())[key].grad), g[f"{tag}.grad.{key}"]) < 2e-4, key
    # routing of the very samples the renderer used is bit-exact with the reference
    from adaptive_city_nerf_b200 import ops
    t = ops.sample_stratified(rays, 32, None)
    pts = ops.points(rays, t)[:, :3].contiguous()
    wgt, hard = m._routing(pts)
    if tag == "soft":
        assert ((npy(wgt) > 0) == g["soft.support"]).all()
    else:
        assert (npy(hard) == g["hard.assign"]).all()


def test_chunking_is_transparent(golden):
    g = golden("render")
    render_rays = _render()
    m = make_container(1, np.zeros((1, 3)), [synth.AABB_GLOBAL], 1.0, False, 83).eval()
    with torch.no_grad():
        a = render_rays(m, cu(g["rays"]), ray_samples=32, active_module=0, chunk=1 << 20)
        b = render_rays(m, cu(g["rays"]), ray_samples=32, active_module=0, chunk=1000)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_render_image_and_bg_policies(golden):
    from adaptive_city_nerf_b200.nerfs.ray_rendering import render_image, get_bg_default_color, apply_bg_mask
    from adaptive_city_nerf_b200.nerfs.scene_box import SceneBox
    m = make_container(1, np.zeros((1, 3)), [synth.AABB_GLOBAL], 1.0, False, 83).eval()
    cam = synth.nadir_rays(7, 1, H=16, W=24, f=20.0)[0]
    box = SceneBox(cu(synth.AABB_GLOBAL))
    rgb, depth, acc = render_image(m, H=16, W=24, fx=cam["fx"], fy=cam["fy"], cx=cam["cx"], cy=cam["cy"],
                                   c2w=cu(cam["c2w"]), scene_box=box, active_module=0, ray_samples=16)
    assert rgb.shape == (16, 24, 3) and depth.shape == (384,) and acc.shape == (384,)
    assert torch.isfinite(rgb).all() and float(rgb.min()) >= 0 and float(rgb.max()) <= 1
    assert get_bg_default_color(None, 3, "none") is None
    assert float(get_bg_default_color(rgb, 3, "black").sum()) == 0
    x = torch.zeros(4, 3, device="cuda")
    apply_bg_mask(x, torch.tensor([True, False, True, False], device="cuda"), "white")
    assert x.sum() == 6


def test_training_psnr_matches_reference(golden):
    """150 Adam steps on an analytic scene from identical init, batches and jitter: PSNR within 0.1 dB
    of the reference's trajectory end (north star), fp32 path."""
    g = golden("train")
    render_rays = _render()
    conf = dict(levels=16, features_per_level=2, log2_hashmap_size=14, max_res=4096, min_res=16, interpolation="Linear")
    m = make_container(1, np.zeros((1, 3)), [synth.AABB_GLOBAL], 1.0, False, 91, hash_conf=conf)
    with torch.no_grad():
        m.submodules[0].xyz_encoder.hash_table.mul_(1e-3 / 0.5)
    groups = m.get_param_groups()
    opt = torch.optim.Adam([
        {"params": groups["encoding"]["params"], "lr": 1e-2},
        {"params": groups["sigma"]["params"], "lr": 2e-3},
        {"params": groups["color"]["params"], "lr": 2e-3}], eps=1e-15)
    all_rays, gt = cu(g["all_rays"]), cu(g["gt"])
    steps, N, S = g["batch_idx"].shape[0], g["batch_idx"].shape[1], 32
    torch.manual_seed(int(g["jitter_seed"]))
    jit = torch.rand(steps, N, S).cuda()            # CPU generator, same stream as the golden run
    m.train()
    psnr = []
    for it in range(steps):
        idx = cu(g["batch_idx"][it]).long()
        rgb, *_ = render_rays(m, all_rays[idx], ray_samples=S, active_module=0, jitter=jit[it])
        loss = torch.nn.functional.mse_loss(rgb, gt[idx])
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        psnr.append(-10.0 * np.log10(float(loss) + 1e-24))
    psnr = np.array(psnr)
    ref = g["psnr"]
    assert abs(psnr[0] - ref[0]) < 1e-3                       # identical start
    assert abs(psnr[-10:].mean() - ref[-10:].mean()) < 0.1    # north-star bar
    m.eval()
    with torch.no_grad():
        rgb, *_ = render_rays(m, all_rays[:2048], ray_samples=S, active_module=0)
    final = -10.0 * np.log10(float(torch.nn.functional.mse_loss(rgb, gt[:2048])) + 1e-24)
    assert abs(final - float(g["final_eval_psnr"])) < 0.1


@pytest.mark.parametrize("loss", ["linear", "srgb", "plain"])
def test_graphed_task_adapt_matches_eager(golden, loss):
    """The inner loop (meta_core.py:14-68, first order) replayed as one CUDA graph gives the weights the launch-by-launch
    loop gives, task after task, and its fast weights drive a first-order outer gradient.  Loss: compute_mse_loss in
    P.color_space (the fused epilogue) or a caller-supplied function."""
    from adaptive_city_nerf_b200.meta import GraphedTaskAdapt, accumulate_first_order_grads, task_adapt_eager
    from adaptive_city_nerf_b200.nerfs.ray_rendering import render_rays
    m = make_container(1, np.zeros((1, 3), F32), [synth.AABB_GLOBAL], 1.0, False, seed0=500).eval()   # eval: no jitter
    ex = m.submodules[0]
    N, S = 1500, 24
    kw = dict(active_module=0, ray_samples=S, iterations=4, inner_lr=5e-2)
    kw.update(dict(loss_fn=torch.nn.functional.mse_loss) if loss == "plain" else dict(color_space=loss))
    adapt = GraphedTaskAdapt(m, n_rays=N, **kw)
    for task in range(3):
        o, d = synth.random_rays_in_box(600 + task, N)
        rays = torch.cat([cu(o), cu(d), torch.zeros(N, 1, device="cuda"), torch.full((N, 1), 0.4, device="cuda")], dim=1)
        rgbs = torch.rand(N, 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(task))
        fast_g, losses_g = adapt(rays, rgbs)
        fast_e, losses_e = task_adapt_eager(m, rays, rgbs, **kw)
        for k in fast_e:
            assert torch.allclose(fast_g[k], fast_e[k].detach(), rtol=0, atol=1e-6), (task, k)
        assert abs(float(losses_g[-1]) - float(losses_e[-1])) < 1e-6
        assert float(losses_e[-1]) < float(losses_e[0])                    # it does adapt
        with torch.no_grad():                                              # an "outer update": theta moves between tasks
            for p in ex.sigma_head.parameters():
                p.add_(0.01)
    fast = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in fast_g.items())
    rgb, *_ = render_rays(m, rays, ray_samples=S, params=fast, active_module=0)
    ((rgb - rgbs) ** 2).mean().backward()
    ex.zero_grad(set_to_none=True)
    accumulate_first_order_grads(ex, fast)
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for _, p in ex.meta_named_parameters())


def test_inner_loop_backward_skips_the_table(golden):
    """torch.autograd.grad(loss, fast weights) -- the inner loop -- must not compute d_enc or scatter into the hash table
    (the fused autograd node learns from the engine what this pass wants); a full backward still does."""
    from adaptive_city_nerf_b200 import _lib
    from adaptive_city_nerf_b200.nerfs.ray_rendering import render_rays
    m = make_container(1, np.zeros((1, 3), F32), [synth.AABB_GLOBAL], 1.0, False, seed0=510).eval()
    ex = m.submodules[0]
    N, S = 700, 16
    o, d = synth.random_rays_in_box(77, N)
    rays = torch.cat([cu(o), cu(d), torch.zeros(N, 1, device="cuda"), torch.full((N, 1), 0.4, device="cuda")], dim=1)
    fast = OrderedDict((n, p) for n, p in ex.meta_named_parameters())
    for amp in (False, True):
        with torch.autocast("cuda", enabled=amp, dtype=torch.float16):
            rgb, *_ = render_rays(m, rays, ray_samples=S, params=fast, active_module=0)
        loss = (rgb ** 2).mean()
        _lib._Profile.start()
        g_inner = torch.autograd.grad(loss, tuple(fast.values()), retain_graph=True)
        torch.cuda.synchronize()
        calls = set(_lib._Profile.stop())
        assert "acn_field_bwd" in calls and not any(k.startswith("acn_hashgrid_bwd") for k in calls), calls
        _lib._Profile.start()
        loss.backward()
        torch.cuda.synchronize()
        calls = set(_lib._Profile.stop())
        assert any(k.startswith("acn_hashgrid_bwd") or k == "acn_render_expert_bwd" for k in calls), calls   # fp32: two kernels; fp16: fused
        assert ex.xyz_encoder.hash_table.grad is not None and float(ex.xyz_encoder.hash_table.grad.abs().sum()) > 0
        for (n, p), g in zip(ex.meta_named_parameters(), g_inner):
            assert torch.allclose(p.grad, g, rtol=1e-4, atol=1e-7), n          # same weight gradients either way
        m.zero_grad(set_to_none=True)


def test_allow_tf32_selects_the_tensor_core_path(golden):
    """Without autocast the strict fp32 kernels run -- unless the caller allowed TF32 matmuls, as the reference's runner
    does (nerf_runner.py:41-42): then the tcgen05 kernels serve the fp32-mode calls too, at TF32-level accuracy."""
    from adaptive_city_nerf_b200 import _lib
    from adaptive_city_nerf_b200.nerfs.ray_rendering import render_rays
    m = make_container(1, np.zeros((1, 3), F32), [synth.AABB_GLOBAL], 1.0, False, seed0=520).eval()
    N, S = 900, 16
    o, d = synth.random_rays_in_box(78, N)
    rays = torch.cat([cu(o), cu(d), torch.zeros(N, 1, device="cuda"), torch.full((N, 1), 0.4, device="cuda")], dim=1)
    prev = torch.backends.cuda.matmul.allow_tf32
    try:
        torch.backends.cuda.matmul.allow_tf32 = False
        rgb32, dep32, *_ = render_rays(m, rays, ray_samples=S, active_module=0)
        (rgb32.sum() + dep32.sum()).backward()
        g32 = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
        m.zero_grad(set_to_none=True)
        torch.backends.cuda.matmul.allow_tf32 = True
        _lib._Profile.start()
        rgb, dep, *_ = render_rays(m, rays, ray_samples=S, active_module=0)
        (rgb.sum() + dep.sum()).backward()
        torch.cuda.synchronize()
        _lib._Profile.stop()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    assert rgb.dtype == torch.float32
    assert (rgb - rgb32).abs().max() < 1e-3 and (dep - dep32).abs().max() < 1e-3
    assert not torch.equal(rgb, rgb32)                                   # a different (tensor-core) kernel did run
    for n, p in m.named_parameters():
        if p.grad is not None:
            a, b = p.grad.double(), g32[n].double()
            assert float((a - b).norm() / (b.norm() + 1e-30)) < 3e-2, n

@pytest.mark.parametrize("margin", [1.0, 1.05])
def test_container_forward_rays_equals_forward_of_points(margin):
    """The container's render path routes and buckets straight from the rays; it must give exactly what
    `model(points(rays, t))` gives (forward bit for bit; gradients up to the order of the table atomics)."""
    from adaptive_city_nerf_b200 import ops
    m = make_container(4, synth.CENTROIDS_G22, synth.EXPERT_BOXES_G22, margin, False, seed0=530).eval()
    N, S = 777, 20
    o, d = synth.random_rays_in_box(81, N)
    rays = torch.cat([cu(o), cu(d), torch.zeros(N, 1, device="cuda"), torch.full((N, 1), 0.5, device="cuda")], dim=1)
    t = ops.sample_stratified(rays, S, None)
    grads = []
    outs = []
    for fused in (True, False):
        m.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.float16):
            y = m.forward_rays(rays, t) if fused else m(ops.points(rays, t)).view(N, S, 4)
        (y[..., :3].sum() + 0.1 * y[..., 3].clamp(max=50).sum()).backward()
        outs.append(y.detach())
        grads.append({n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None})
    assert torch.equal(outs[0], outs[1])
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):       # frames order the buckets ray-major
        assert torch.equal(m.forward_rays(rays, t, ray_major=True), outs[0])
    assert grads[0].keys() == grads[1].keys() and len(grads[0]) > 0
    for n in grads[0]:
        a, b = grads[0][n].double(), grads[1][n].double()
        assert float((a - b).norm() / (b.norm() + 1e-30)) < 1e-4, n

def test_coherent_frames_are_detected_and_render_identically():
    """Consecutive rays of a frame are adjacent pixels: render_rays finds that out (inference only) and switches the
    gather kernels to one-sample-of-32-rays warps; shuffled rays do not trigger it; the image is the same bit for bit."""
    from adaptive_city_nerf_b200 import _lib, ops
    from adaptive_city_nerf_b200.nerfs.ray_rendering import render_rays
    from adaptive_city_nerf_b200.nerfs.ray_sampling import clamp_rays_near_far, get_ray_directions, get_rays
    from adaptive_city_nerf_b200.nerfs.scene_box import SceneBox
    box = SceneBox(cu(synth.AABB_GLOBAL))
    cam = synth.nadir_rays(9, 1, H=40, W=72, f=900.0)[0]
    dirs = get_ray_directions(40, 72, cam["fx"], cam["fy"], cam["cx"], cam["cy"], True, torch.device("cuda"))
    rays, _ = clamp_rays_near_far(get_rays(dirs, cu(cam["c2w"]), scene_box=box).view(-1, 8), (None, None))
    shuffled = rays[torch.randperm(rays.shape[0], device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))]
    assert ops.rays_are_coherent(rays, 32) and not ops.rays_are_coherent(shuffled, 32)
    assert not ops.rays_are_coherent(rays[:20], 32)
    for K in (1, 4):
        if K == 1:
            m = make_container(1, np.zeros((1, 3), F32), [synth.AABB_GLOBAL], 1.0, False, seed0=540).eval()
        else:
            m = make_container(4, synth.CENTROIDS_G22, synth.EXPERT_BOXES_G22, 1.05, True, seed0=541).eval()
        kw = dict(ray_samples=32, active_module=0 if K == 1 else None)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            auto = render_rays(m, rays, **kw)
            plain = render_rays(m, rays, coherent_rays=False, **kw)
            forced = render_rays(m, rays, coherent_rays=True, **kw)
        for a, b, c in zip(auto, plain, forced):
            assert torch.equal(a, b) and torch.equal(a, c)
