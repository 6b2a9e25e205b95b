"""acn_render_expert_fwd (hash encode + tcgen05 MLPs in one warp-specialised kernel) against the two-kernel path
acn_hashgrid_fwd[_rays](fp16) -> acn_field_fwd(ACN_F16): the producers run the encode kernel's arithmetic, so the fp16
encoding rows are bit-identical; the first MLP layer reads them from shared memory instead of tensor memory (same fp16
products, fp32 accumulation), so [rgb, sigma] agrees to accumulation order."""
import numpy as np
import pytest
import torch

import synth
from helpers import F32, cu, npy

pytestmark = pytest.mark.gpu


def _setup(L, log2T, interp, N, S, seed, G=15):
    from adaptive_city_nerf_b200 import ops, _lib
    E = 2 * L
    sd = synth.make_expert_params(17, E=E, G=G, log2T=4)
    wt = [cu(w) for w in synth.expert_weight_list(sd)]
    gen = torch.Generator(device="cuda").manual_seed(seed)
    res_np = np.floor(16 * np.exp(np.arange(L) * np.log(4096 / 16) / max(L - 1, 1))).astype(np.int32)
    spec = ops.GridSpec(L, 2, log2T, torch.tensor(res_np), _lib.INTERP[interp], res_host=res_np.tolist())
    table = (torch.rand(L << log2T, 2, device="cuda", generator=gen) - 0.5) * 0.2
    lo, hi = synth.AABB_GLOBAL
    box6 = cu(np.concatenate([lo, hi - lo]).astype(F32))
    o = torch.rand(N, 3, device="cuda", generator=gen) * cu(hi - lo) * 0.5 + cu(lo) + cu(hi - lo) * 0.1
    d = torch.nn.functional.normalize(torch.randn(N, 3, device="cuda", generator=gen), dim=-1)
    rays = torch.cat([o, d, torch.zeros(N, 1, device="cuda"), torch.ones(N, 1, device="cuda")], 1).contiguous()
    t = (torch.rand(N, S, device="cuda", generator=gen) * 0.6).sort(dim=1).values.contiguous()
    return ops, spec, table, box6, rays, t, wt


def _close(y, y_ref):
    assert torch.isfinite(y).all()
    assert float((y[:, :3] - y_ref[:, :3]).abs().max()) < 1e-6
    assert float(((y[:, 3] - y_ref[:, 3]).abs() / (y_ref[:, 3].abs() + 1e-6)).max()) < 1e-5


@pytest.mark.parametrize("L,log2T,interp,N,S,G", [(16, 14, "Linear", 1400, 48, 15), (16, 12, "Smoothstep", 301, 49, 15),
                                                  (8, 10, "Linear", 517, 33, 7), (16, 19, "Linear", 2100, 64, 15)])
def test_fused_forward_from_rays(L, log2T, interp, N, S, G):
    ops, spec, table, box6, rays, t, wt = _setup(L, log2T, interp, N, S, seed=L + log2T + N, G=G)
    enc_ref = ops.hashgrid_fwd_rays(rays, t, table, spec, box6, torch.float16)
    y_ref = ops.field_fwd(enc_ref, rays[:, 3:], 8, S, wt, half=True)
    for stage in (True, False):
        ops.STAGE_COARSE_LEVELS = stage
        try:
            y, enc = ops.render_expert_fwd((rays, t), table, spec, box6, rays[:, 3:], 8, S, wt, want_enc=True)
        finally:
            ops.STAGE_COARSE_LEVELS = False
        assert torch.equal(enc, enc_ref), f"staged={stage}: {int((enc != enc_ref).sum())} encoding values differ"
        _close(y, y_ref)
    # inference: no encoding written; ray-major tile order (frames) gives every point the same value
    y2, enc2 = ops.render_expert_fwd((rays, t), table, spec, box6, rays[:, 3:], 8, S, wt, want_enc=False)
    assert enc2 is None and torch.equal(y2, y)
    y3, enc3 = ops.render_expert_fwd((rays, t), table, spec, box6, rays[:, 3:], 8, S, wt, want_enc=True, ray_major=True)
    assert torch.equal(enc3, enc_ref)
    _close(y3, y_ref)
    flag = torch.ones(1, dtype=torch.int32, device="cuda")
    y4, _ = ops.render_expert_fwd((rays, t), table, spec, box6, rays[:, 3:], 8, S, wt, want_enc=False, ray_major=flag)
    assert torch.equal(y4, y3)


@pytest.mark.parametrize("P", [1, 127, 129, 70_001])
def test_fused_forward_from_points_and_row_ranges(P):
    ops, spec, table, box6, rays, t, wt = _setup(16, 13, "Linear", 1500, 48, seed=P)
    x6 = ops.points(rays, t)[:P].contiguous()
    enc_ref = ops.hashgrid_fwd(x6, table, spec, box6, torch.float16)
    y_ref = ops.field_fwd(enc_ref, x6[:, 3:], 6, 1, wt, half=True)
    y, enc = ops.render_expert_fwd((x6,), table, spec, box6, x6[:, 3:], 6, 1, wt, want_enc=True)
    assert torch.equal(enc, enc_ref)
    _close(y, y_ref)
    # no box (positions already in the unit cube): coarse levels are not staged, values unchanged
    u6 = x6.clone()
    u6[:, :3] = torch.rand(P, 3, device="cuda")
    u6[0, :3] = torch.tensor([1.0, 0.0, 0.5])                         # on the boundary: corner rows beyond the lattice
    e0 = ops.hashgrid_fwd(u6, table, spec, None, torch.float16)
    y0, e1 = ops.render_expert_fwd((u6,), table, spec, None, u6[:, 3:], 6, 1, wt, want_enc=True)
    assert torch.equal(e1, e0)
    _close(y0, ops.field_fwd(e0, u6[:, 3:], 6, 1, wt, half=True))
    if P > 1000:      # a bucket: rows [first, end) known only to the device; rows outside are never written
        rng = torch.tensor([1234, 66_001], dtype=torch.int32, device="cuda")
        out = torch.full((P, 4), -7.0, device="cuda")
        encb = torch.full((P, 32), -7.0, dtype=torch.float16, device="cuda")
        ops.render_expert_fwd((x6,), table, spec, box6, x6[:, 3:], 6, 1, wt, want_enc=True, rng=rng, enc=encb, out=out)
        assert torch.equal(encb[1234:66_001], enc_ref[1234:66_001]) and bool((encb[:1234] == -7).all()) and bool((encb[66_001:] == -7).all())
        _close(out[1234:66_001], y_ref[1234:66_001])
        assert bool((out[:1234] == -7).all()) and bool((out[66_001:] == -7).all())
        empty = torch.tensor([5, 5], dtype=torch.int32, device="cuda")
        out2 = torch.full((P, 4), -7.0, device="cuda")
        ops.render_expert_fwd((x6,), table, spec, box6, x6[:, 3:], 6, 1, wt, want_enc=False, rng=empty, out=out2)
        assert bool((out2 == -7).all())


def test_render_rays_same_image_and_gradients_with_either_forward():
    """The public path: render_rays under autocast with the fused forward on / off -- same rgb, depth, and (training) the
    same table and MLP gradients, since the backward consumes bit-identical encodings."""
    from adaptive_city_nerf_b200 import ops
    from adaptive_city_nerf_b200.nerfs.ray_rendering import render_rays
    from helpers import make_container
    lo, hi = synth.AABB_GLOBAL
    m = make_container(1, [[0.0, 0.0]], [synth.AABB_GLOBAL], 1.0, False, 40, hash_conf=dict(levels=16, features_per_level=2,
                       log2_hashmap_size=15, max_res=4096, min_res=16, interpolation="Linear"))
    gen = torch.Generator(device="cuda").manual_seed(3)
    N = 3000
    o = torch.rand(N, 3, device="cuda", generator=gen) * cu(hi - lo) * 0.3 + cu(lo) + cu(hi - lo) * 0.3
    d = torch.nn.functional.normalize(torch.randn(N, 3, device="cuda", generator=gen), dim=-1)
    rays = torch.cat([o, d, torch.full((N, 1), 0.01, device="cuda"), torch.full((N, 1), 0.9, device="cuda")], 1)
    res = {}
    for fused in (True, False):
        ops.FUSED_EXPERT_FWD = fused
        try:
            m.zero_grad(set_to_none=True)
            m.eval()
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                ev = render_rays(m, rays, ray_samples=32, active_module=0)
            m.train()
            torch.manual_seed(5)
            with torch.autocast("cuda", dtype=torch.float16):
                tr = render_rays(m, rays, ray_samples=32, active_module=0)
            (tr[0].float().square().mean() + tr[1].float().mean() * 0.1).backward()
            res[fused] = (ev[0].float(), ev[1].float(), tr[0].float().detach(),
                          {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None})
        finally:
            ops.FUSED_EXPERT_FWD = True
    a, b = res[True], res[False]
    for i in range(3):
        assert float((a[i] - b[i]).abs().max()) < 2e-6
    assert a[3].keys() == b[3].keys() and len(a[3]) >= 15
    for n in a[3]:
        den = float(b[3][n].norm()) + 1e-30
        assert float((a[3][n] - b[3][n]).norm()) / den < 1e-4, n


@pytest.mark.parametrize("N,Hb", [(1, 32), (1000, 32), (70_001, 32), (333, 64), (257, 8)])
def test_background_head_kernel_matches_the_torch_layers(N, Hb):
    """acn_background_fwd/_bwd against the reference's own op sequence (meta_container.py:347-382: F.normalize ->
    SHEncoder -> Linear -> ReLU -> Linear -> Sigmoid) run as fp32 torch layers: rgb 2e-6, weight gradients 1e-5."""
    from adaptive_city_nerf_b200 import ops
    from adaptive_city_nerf_b200.models.encodings import SHEncoder
    gen = torch.Generator(device="cuda").manual_seed(N + Hb)
    rays = torch.randn(N, 8, device="cuda", generator=gen)
    d = rays[:, 3:6]                                   # a strided view, as render_rays passes it
    mlp = torch.nn.Sequential(torch.nn.Linear(16, Hb), torch.nn.ReLU(), torch.nn.Linear(Hb, 3), torch.nn.Sigmoid()).cuda()
    with torch.no_grad():
        for p in mlp.parameters():
            p.copy_(torch.randn(p.shape, device="cuda", generator=gen) * 0.5)
    sh = SHEncoder(levels=4)
    gy = torch.randn(N, 3, device="cuda", generator=gen)
    ref = mlp(sh(torch.nn.functional.normalize(d, dim=-1)))
    (ref * gy).sum().backward()
    gref = [p.grad.clone() for p in mlp.parameters()]
    for p in mlp.parameters():
        p.grad = None
    out = ops.BackgroundFn.apply(d, mlp[0].weight, mlp[0].bias, mlp[2].weight, mlp[2].bias, False)
    assert out.dtype == torch.float32 and float((out - ref).abs().max()) < 2e-6
    (out * gy).sum().backward()
    for p, gr in zip(mlp.parameters(), gref):
        assert float((p.grad - gr).norm()) <= 1e-5 * float(gr.norm()) + 1e-7
    half = ops.BackgroundFn.apply(d, mlp[0].weight, mlp[0].bias, mlp[2].weight, mlp[2].bias, True)
    assert half.dtype == torch.float16 and float((half.float() - ref).abs().max()) < 6e-4


def test_many_training_steps_queued_back_to_back_do_not_fault():
    """120 training steps launched without a host sync in between (what bench.py and a real training loop do): the
    warp-specialised kernels' mbarrier waits are bounded by wall time, not by a poll count -- a poll-count bound faulted about
    once per hundred queued steps (never with a sync after every step).  Loss finite and decreasing at the end."""
    import bench
    from adaptive_city_nerf_b200.nerfs.losses import mse_in_color_space
    from adaptive_city_nerf_b200.nerfs.ray_rendering import render_rays
    from adaptive_city_nerf_b200.optim import FusedAdam
    dev = torch.device("cuda")
    rays, gt, box = bench.gpu_workload(dev, 100)
    rays, gt = rays[: 1 << 16].contiguous(), gt[: 1 << 16].contiguous()
    model = bench.make_model(dev, box)
    opt = FusedAdam([{"params": list(model.parameters()), "lr": 2e-3}], eps=1e-15)
    losses = []
    for _ in range(120):
        with torch.autocast("cuda", dtype=torch.float16):
            rgb, *_ = render_rays(model, rays, ray_samples=bench.SAMPLES, active_module=0, chunk=1 << 30)
        loss = mse_in_color_space(rgb, gt, "linear")
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step(max_norm=1.0)
        losses.append(loss.detach())
    torch.cuda.synchronize()
    ls = torch.stack(losses).cpu()
    assert bool(torch.isfinite(ls).all()) and float(ls[-10:].mean()) < float(ls[:10].mean())
