"""GPU parity tests for what surrounds the render in a training step: the colour-space + MSE
loss epilogue (`acn_color_mse`) and the fused unscale / clip / Adam tail (`acn_grad_sqnorm`, `acn_adam_prepare`,
`acn_adam_apply`) -- against the CPU oracle on the same seeded inputs, against the golden fixtures generated from
the reference (`tests/golden/{loss,optim}.npz`), and against torch's own optimizers on the device.

Tolerances: squared errors 5e-7 abs (powf differs from torch.pow by an ulp; values in [0,1]); loss 2e-6 rel;
gradients 2e-5 rel; parameters after an optimizer step 2e-6 rel + 2e-7 abs (same op order as torch.optim.Adam, constants
rounded from the same doubles)."""
import numpy as np
import pytest
import torch

import synth
from helpers import F32, assert_bitexact, cu, npy

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from adaptive_city_nerf_b200 import ops as o
    return o


# ----------------------------------------------------------------------------- loss epilogue
@pytest.mark.parametrize("cs", ["linear", "srgb", "identity"])
def test_color_mse_vs_oracle_and_reference(ops, orc, golden, cs):
    g = golden("loss")
    pred, gt = synth.loss_inputs()
    loss, dpred = ops.color_mse(cu(pred), cu(gt), cs, "mean", want_grad=True)
    elem, dsum = ops.color_mse(cu(pred), cu(gt), cs, "none", want_grad=True)
    o_loss, o_dpred = orc.color_mse(pred, gt, cs, "mean")
    o_elem, o_dsum = orc.color_mse(pred, gt, cs, "none")
    np.testing.assert_allclose(npy(elem), o_elem, rtol=0, atol=5e-7)
    np.testing.assert_allclose(npy(elem), g[f"{cs}_elem"], rtol=0, atol=5e-7)
    np.testing.assert_allclose(float(loss), float(o_loss), rtol=2e-6)
    np.testing.assert_allclose(float(loss), float(g[f"{cs}_loss"]), rtol=2e-6)
    np.testing.assert_allclose(npy(dpred), o_dpred, rtol=2e-5, atol=1e-9)
    np.testing.assert_allclose(npy(dsum), o_dsum, rtol=2e-5, atol=1e-7)
    ref = g[f"{cs}_grad"]
    ok = ~np.isnan(ref)                     # sRGB: the reference is NaN at pred == 0; checked against the oracle above
    assert ok.sum() >= ref.size - 2
    np.testing.assert_allclose(npy(dpred)[ok], ref[ok], rtol=2e-5, atol=1e-9)
    assert np.isfinite(npy(dpred)).all()
    total, _ = ops.color_mse(cu(pred), cu(gt), cs, "sum")
    np.testing.assert_allclose(float(total), float(o_elem.astype(np.float64).sum()), rtol=2e-6)


def _torch_color_mse(pred, gt, cs):
    """The reference's formulas (nerfs/color_space.py) restated in torch on the device, for autograd comparison."""
    gt = gt.clamp(0, 1)
    if cs == "linear":
        gt = torch.where(gt <= 0.04045, gt / 12.92, ((gt + 0.055) / 1.055).pow(2.4)).clamp(0, 1)
        pred = pred.clamp(0, 1)
    elif cs == "srgb":
        x = pred.clamp(0, 1)
        pred = torch.where(x <= 0.0031308, 12.92 * x, 1.055 * x.pow(1 / 2.4) - 0.055).clamp(0, 1)
    return pred, gt


@pytest.mark.parametrize("cs", ["linear", "srgb", "identity"])
@pytest.mark.parametrize("reduction", ["mean", "none"])
def test_color_mse_autograd_and_large_batch(ops, cs, reduction):
    """2^18 rays (BASELINE configs[1] batch): the fused node's forward and backward vs torch autograd over the same
    formulas, upstream gradient = a GradScaler-like scale; the reduction is bit-reproducible across launches."""
    from adaptive_city_nerf_b200.nerfs.losses import mse_in_color_space
    gen = torch.Generator(device="cuda").manual_seed(5)
    N = 1 << 18
    pred = (torch.rand(N, 3, device="cuda", generator=gen) * 1.2 - 0.1)
    pred[pred == 0] = 0.5                                         # keep the reference's NaN point out of this comparison
    gt = torch.rand(N, 3, device="cuda", generator=gen)
    p1 = pred.clone().requires_grad_()
    out = mse_in_color_space(p1, gt, cs, reduction)
    up = torch.full_like(out, 1024.0) if reduction == "none" else torch.tensor(1024.0, device="cuda")
    out.backward(up)
    p2 = pred.clone().requires_grad_()
    a, b = _torch_color_mse(p2, gt, cs)
    ref = torch.nn.functional.mse_loss(a, b, reduction=reduction)
    ref.backward(up)
    if reduction == "mean":
        np.testing.assert_allclose(float(out.detach()), float(ref.detach()), rtol=2e-6)
        again = mse_in_color_space(pred, gt, cs, reduction)
        assert float(again) == float(out.detach())                        # fixed-order reduction
    else:
        np.testing.assert_allclose(npy(out), npy(ref), rtol=0, atol=1e-6)
    # |pred' - gt'| can cancel to a few ulp, so the absolute bar is 2 ulp(0.5) x the largest slope (12.92) x upstream
    atol = 2 * 6e-8 * 2 * 12.92 * 1024 * (1.0 / pred.numel() if reduction == "mean" else 1.0)
    np.testing.assert_allclose(npy(p1.grad), npy(p2.grad), rtol=3e-5, atol=atol)


def test_color_mse_edge_cases(ops):
    empty = torch.empty(0, 3, device="cuda")
    loss, _ = ops.color_mse(empty, empty, "linear", "mean")
    assert np.isnan(float(loss))                                  # torch: mean over nothing = nan
    total, _ = ops.color_mse(empty, empty, "linear", "sum")
    assert float(total) == 0.0
    one = torch.tensor([[0.25, float("nan"), 2.0]], device="cuda")
    elem, d = ops.color_mse(one, torch.tensor([[0.5, 0.5, 1.0]], device="cuda"), "linear", "none", want_grad=True)
    e = npy(elem)[0]
    assert np.isnan(e[1]) and e[2] == 0.0 and npy(d)[0, 2] == 0.0  # NaN propagates like torch.clamp; clamped side has no gradient
    with pytest.raises(ValueError):
        ops.color_mse(one, one, "xyz")
    with pytest.raises(RuntimeError):
        ops.color_mse(one.cpu(), one.cpu(), "linear")             # no CPU path
    gt_b = torch.tensor([0.2, 0.4, 0.6], device="cuda")           # broadcast ground truth (a constant background)
    l1, _ = ops.color_mse(one.nan_to_num(0.3), gt_b, "identity")
    l2 = torch.nn.functional.mse_loss(one.nan_to_num(0.3), gt_b.expand(1, 3))
    np.testing.assert_allclose(float(l1), float(l2), rtol=1e-6)


# ----------------------------------------------------------------------------- optimizer tail
def _fixture_params():
    params, grads = synth.optim_inputs()
    ps = [torch.nn.Parameter(cu(a).clone()) for a in params]
    groups = {}
    for (grp, _), p in zip(synth.OPTIM_SHAPES, ps):
        groups.setdefault(grp, []).append(p)
    return ps, grads, [{"params": v, "lr": synth.OPTIM_LRS[k], "name": k} for k, v in groups.items()]


def _check_step(ps, g, name, it):
    for k, p in enumerate(ps):
        np.testing.assert_allclose(npy(p), g[f"{name}_p{k}_step{it}"], rtol=2e-6, atol=2e-7, err_msg=f"{name} step {it} tensor {k}")


@pytest.mark.parametrize("flow", ["step_scaled", "scaler_protocol", "reference_calls"])
@pytest.mark.parametrize("name,adamw,wd", [("adam", False, 0.0), ("adamw", True, 0.05), ("adam_wd", False, 0.05)])
def test_fused_adam_vs_reference_meta_update(golden, name, adamw, wd, flow):
    """Six steps of maml_meta_update (GradScaler at 65536, clip 1.0, one overflowing gradient) -- parameters, moments,
    skipped step and loss-scale trajectory against the fixture produced by the reference's own functions."""
    from adaptive_city_nerf_b200.optim import FusedAdam
    g = golden("optim")
    ps, grads, groups = _fixture_params()
    opt = FusedAdam(groups, lr=1e-3, weight_decay=wd, adamw=adamw)
    scaler = torch.amp.GradScaler("cuda", init_scale=65536.0, growth_interval=2)
    scales = []
    for it, gs in enumerate(grads):
        scales.append(scaler.get_scale())
        loss = sum((p * cu(a)).sum() for p, a in zip(ps, gs))
        opt.zero_grad(set_to_none=True)
        scaler.scale(loss).backward()
        if flow == "step_scaled":
            opt.step_scaled(scaler, max_norm=1.0)
        elif flow == "scaler_protocol":
            scaler.step(opt, max_norm=1.0)
            scaler.update()
        else:                                                   # the reference's literal sequence (meta_core.py:131-137)
            scaler.unscale_(opt)
            torch.nn.utils.clip_grad_norm_(ps, 1.0)
            scaler.step(opt)
            scaler.update()
        if it in synth.OPTIM_KEEP:
            _check_step(ps, g, name, it)
    scales.append(scaler.get_scale())
    assert scales == list(g[f"{name}_scales"])
    assert float(opt.steps_taken) == float(g[f"{name}_steps"]) == 5.0
    if name == "adam":
        for k, p in enumerate(ps):
            np.testing.assert_allclose(npy(opt.state[p]["exp_avg"]), g[f"adam_m{k}"], rtol=2e-6, atol=1e-9)
            np.testing.assert_allclose(npy(opt.state[p]["exp_avg_sq"]), g[f"adam_v{k}"], rtol=2e-6, atol=1e-12)


def test_fused_adam_vs_oracle_fp32_flow(orc):
    """No scaler: clip at 0.5 then Adam, against the CPU oracle; `last_norm` is clip_grad_norm_'s return value."""
    from adaptive_city_nerf_b200.optim import FusedAdam
    params, grads = synth.optim_inputs(seed=71, steps=4)
    lrs = [synth.OPTIM_LRS[grp] for grp, _ in synth.OPTIM_SHAPES]
    st = orc.AdamState(params, lrs)
    ps = [torch.nn.Parameter(cu(a).clone()) for a in params]
    opt = FusedAdam([{"params": [p], "lr": lr} for p, lr in zip(ps, lrs)], max_norm=0.5, write_grads=True)
    for it, gs in enumerate(grads):
        if it == 3:
            continue                                            # the overflowing one is for the scaler tests
        norm, skip = st.update(gs, max_norm=0.5)
        for p, a in zip(ps, gs):
            p.grad = cu(a).clone()
        opt.step()
        np.testing.assert_allclose(float(opt.last_norm), norm, rtol=2e-6)
        for k, p in enumerate(ps):
            np.testing.assert_allclose(npy(p), st.p[k], rtol=2e-6, atol=2e-7)
            np.testing.assert_allclose(npy(p.grad), st.g[k], rtol=2e-6, atol=1e-9)   # write_grads: the clipped gradient


def test_fused_adam_table_sized_vs_torch():
    """One expert's worth of parameters (64 MiB table + 14 MLP tensors) against torch.optim.Adam + clip_grad_norm_ on
    the device, and a non-finite gradient leaves every tensor untouched."""
    from adaptive_city_nerf_b200.optim import FusedAdam
    gen = torch.Generator(device="cuda").manual_seed(9)
    shapes = [(16 << 19, 2)] + synth.expert_shapes()
    mk = lambda: [torch.nn.Parameter((torch.rand(s, device="cuda", generator=torch.Generator(device="cuda").manual_seed(i)) - 0.5))
                  for i, s in enumerate(shapes)]
    a, b = mk(), mk()
    ref = torch.optim.Adam([{"params": a[:1], "lr": 1e-2}, {"params": a[1:], "lr": 2e-3}], eps=1e-15)
    ours = FusedAdam([{"params": b[:1], "lr": 1e-2}, {"params": b[1:], "lr": 2e-3}], eps=1e-15, max_norm=1.0)
    for it in range(3):
        gs = [torch.randn(s, device="cuda", generator=gen) * 1e-3 for s in shapes]
        for p, q, g_ in zip(a, b, gs):
            p.grad, q.grad = g_.clone(), g_.clone()
        n_ref = torch.nn.utils.clip_grad_norm_(a, 1.0)
        ref.step()
        ours.step()
        np.testing.assert_allclose(float(ours.last_norm), float(n_ref), rtol=1e-5)
    for p, q in zip(a, b):
        assert torch.allclose(p, q, rtol=2e-6, atol=2e-7), float((p - q).abs().max())
    before = [q.detach().clone() for q in b]
    for q in b:
        q.grad = torch.zeros_like(q)
    b[5].grad[0] = float("inf")
    ours.step()
    assert float(ours.steps_taken) == 3.0
    for q, q0 in zip(b, before):
        assert torch.equal(q, q0)


def test_fused_adam_many_tensors_and_checkpoint():
    """More tensors than one launch carries (8 experts x 15 + background), and a state_dict round trip in
    torch.optim.Adam's layout (the reference checkpoints the optimizer, utils.py save_checkpoint)."""
    from adaptive_city_nerf_b200.optim import FusedAdam
    gen = torch.Generator(device="cuda").manual_seed(3)
    shapes = [(37,), (5, 9), (1,), (128, 3)] * 31                  # 124 tensors
    mk = lambda: [torch.nn.Parameter(torch.full(s, 0.25, device="cuda")) for s in shapes]
    a, b, c = mk(), mk(), mk()
    ref = torch.optim.Adam(a, lr=3e-3)
    ours = FusedAdam(b, lr=3e-3)
    grads = [[torch.randn(s, device="cuda", generator=gen) for s in shapes] for _ in range(4)]
    for it in range(2):
        for p, q, g_ in zip(a, b, grads[it]):
            p.grad, q.grad = g_.clone(), g_.clone()
        ref.step(); ours.step()
    import copy
    sd = copy.deepcopy(ours.state_dict())                        # load_state_dict aliases same-device tensors, as torch's does
    assert float(sd["state"][0]["step"]) == 2.0
    resumed = FusedAdam(c, lr=3e-3)
    with torch.no_grad():
        for q, r in zip(b, c):
            r.copy_(q)
    resumed.load_state_dict(sd)
    as_torch = torch.optim.Adam(mk(), lr=3e-3)
    as_torch.load_state_dict(sd)                                 # the layout torch.optim.Adam expects
    for it in range(2, 4):
        for p, q, r, g_ in zip(a, b, c, grads[it]):
            p.grad, q.grad, r.grad = g_.clone(), g_.clone(), g_.clone()
        ref.step(); ours.step(); resumed.step()
    for p, q, r in zip(a, b, c):
        assert torch.allclose(p, q, rtol=2e-6, atol=2e-7)
        assert torch.equal(q, r)


def test_fused_adam_in_a_cuda_graph():
    """Nothing in the step touches the host, so it captures: replaying the graph == stepping eagerly."""
    from adaptive_city_nerf_b200.optim import FusedAdam
    shapes = [(1000, 2), (64, 32), (3,)]
    mk = lambda: [torch.nn.Parameter(torch.full(s, 0.1, device="cuda")) for s in shapes]
    a, b = mk(), mk()
    static_g = [torch.zeros(s, device="cuda") for s in shapes]
    for p, q, g_ in zip(a, b, static_g):
        p.grad, q.grad = g_, g_
    eager, graphed = FusedAdam(a, lr=1e-2, max_norm=1.0), FusedAdam(b, lr=1e-2, max_norm=1.0)
    gen = torch.Generator(device="cuda").manual_seed(1)
    seq = [[torch.randn(s, device="cuda", generator=gen) for s in shapes] for _ in range(4)]
    for g_, s0 in zip(static_g, seq[0]):
        g_.copy_(s0)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        graphed.step()                                           # allocates the moments outside the capture
    torch.cuda.current_stream().wait_stream(side)
    eager.step()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        graphed.step()
    for it in range(1, 4):
        for g_, s_ in zip(static_g, seq[it]):
            g_.copy_(s_)
        eager.step()
        graph.replay()
    torch.cuda.synchronize()
    for p, q in zip(a, b):
        assert torch.equal(p, q)
    assert float(graphed.steps_taken) == 4.0


def test_get_optimizer_groups_and_training_step():
    """common/utils.py get_optimizer on our container, then a short AMP training loop through compute_mse_loss and the
    fused tail: the loss goes down and matches the same loop driven by torch.optim.Adam + the reference's calls."""
    import types
    from helpers import make_container
    from adaptive_city_nerf_b200.optim import get_optimizer
    from adaptive_city_nerf_b200.nerfs.losses import compute_mse_loss
    N = 2048
    o, d = synth.random_rays_in_box(21, N)
    r = torch.cat([cu(o), cu(d), torch.zeros(N, 1, device="cuda"), torch.full((N, 1), 0.4, device="cuda")], dim=1)
    gt = torch.rand(N, 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    P = types.SimpleNamespace(lr=1e-3, encoding_lr=1e-2, sigma_lr=2e-3, color_lr=2e-3, bg_lr=1e-3, optimizer="adam",
                              weight_decay=0.0, ray_samples=32, chunk_points=1 << 22, color_space="linear")
    losses = {}
    for which in ("fused", "torch"):
        model = make_container(1, np.zeros((1, 3), F32), [synth.AABB_GLOBAL], 1.0, False, seed0=500)
        if which == "fused":
            opt = get_optimizer(P, model)
            assert [g_["name"] for g_ in opt.param_groups] == ["encoding", "sigma", "color"]
            assert [g_["lr"] for g_ in opt.param_groups] == [1e-2, 2e-3, 2e-3]
        else:
            groups = model.get_param_groups()
            opt = torch.optim.Adam([{"params": list(groups[k]["params"]), "lr": lr}
                                    for k, lr in (("encoding", 1e-2), ("sigma", 2e-3), ("color", 2e-3))])
        scaler = torch.amp.GradScaler("cuda")
        model.eval()                                             # no jitter: both loops see the same samples
        hist = []
        for it in range(8):
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.float16):
                loss = compute_mse_loss(P, model, {"rays": r, "rgbs": gt}, active_module=0)
            scaler.scale(loss).backward()
            if which == "fused":
                opt.step_scaled(scaler, max_norm=1.0)
            else:
                scaler.unscale_(opt)
                torch.nn.utils.clip_grad_norm_([p for g_ in opt.param_groups for p in g_["params"]], 1.0)
                scaler.step(opt)
                scaler.update()
            hist.append(float(loss.detach()))
        losses[which] = hist
    assert losses["fused"][-1] < losses["fused"][0]
    np.testing.assert_allclose(losses["fused"], losses["torch"], rtol=2e-3)


# ----------------------------------------------------------------------------- task-grid binning (ray-batch producer)
@pytest.mark.parametrize("tag,cells,region", [("auto", (1, 6, 6), None), ("box", (2, 3, 4), "expert0")])
def test_dda_task_binning(orc, golden, tag, cells, region):
    """TaskDataset._route_and_bin, "dda" policy: cell per ray, counts and bins -- bit-exact vs the reference's fixture and
    the oracle; the host-side grid tensors equal the reference's."""
    from adaptive_city_nerf_b200.data import TaskGrid, route_and_bin
    from adaptive_city_nerf_b200.data.task_binning import dda_route_rays
    g = golden("taskgrid")
    rays_np = synth.task_rays()
    rays = cu(rays_np)
    reg = None if region is None else tuple(map(tuple, synth.EXPERT_BOXES_G22[0].tolist()))
    grid = TaskGrid(rays, cells, reg)
    np.testing.assert_allclose(npy(grid.aabb), g[f"{tag}_aabb"], rtol=0, atol=0)
    np.testing.assert_allclose(npy(grid.cell_bounds), g[f"{tag}_cell_bounds"], rtol=0, atol=1e-7)
    cid, counts, blen = dda_route_rays(rays, grid, want_len=True)
    o_cid, o_len, o_counts = orc.dda_route_rays(rays_np, npy(grid.aabb), cells, npy(grid.cell3), npy(grid.cell_bounds), npy(grid.tol))
    assert (cid.cpu().numpy() == o_cid).all() and (counts.cpu().numpy() == o_counts).all()
    assert_bitexact(npy(blen), o_len, "in-cell length vs oracle")
    assert (cid.cpu().numpy() == g[f"{tag}_cid"]).all(), int((cid.cpu().numpy() != g[f"{tag}_cid"]).sum())
    assert (counts.cpu().numpy() == g[f"{tag}_counts"]).all()
    assert_bitexact(npy(blen), g[f"{tag}_best_len"], "in-cell length vs reference")
    bins, _ = route_and_bin(rays, cells, reg)
    assert len(bins) == int(np.prod(cells))
    for c, b in enumerate(bins):
        assert b.dtype == torch.int64
        assert sorted(b.tolist()) == np.nonzero(g[f"{tag}_cid"] == c)[0].tolist()


def test_dda_task_binning_large_and_edge_cases(orc):
    """An expert's worth of rays (2^20) against the oracle; empty input; a grid the rays never reach."""
    from adaptive_city_nerf_b200.data import TaskGrid, route_and_bin
    from adaptive_city_nerf_b200.data.task_binning import dda_route_rays
    base = synth.task_rays(seed=91, n_soup=60000)
    rng = np.random.default_rng(2)
    rays_np = base[rng.integers(0, base.shape[0], 1 << 20)]
    rays_np[:, :3] += rng.normal(0, 1e-3, (rays_np.shape[0], 3)).astype(F32)
    rays = cu(rays_np)
    grid = TaskGrid(rays, (1, 12, 12), tuple(map(tuple, synth.AABB_GLOBAL.tolist())))
    cid, counts = dda_route_rays(rays, grid)
    o_cid, _, o_counts = orc.dda_route_rays(rays_np, npy(grid.aabb), grid.cells, npy(grid.cell3), npy(grid.cell_bounds), npy(grid.tol))
    assert (cid.cpu().numpy() == o_cid).all() and (counts.cpu().numpy() == o_counts).all()
    assert int(counts.sum()) > (1 << 19)
    bins, _ = route_and_bin(rays[:0], (1, 2, 2), tuple(map(tuple, synth.AABB_GLOBAL.tolist())))
    assert [b.numel() for b in bins] == [0, 0, 0, 0]
    far_away = ((10.0, 10.0, 10.0), (11.0, 11.0, 11.0))
    bins, _ = route_and_bin(rays[:5000], (1, 3, 3), far_away)
    assert sum(b.numel() for b in bins) == 0
    with pytest.raises(RuntimeError, match="no CPU path"):
        route_and_bin(rays[:10].cpu(), (1, 2, 2))
