"""Pins the CPU oracle (oracle/acn_oracle.c) against outputs of the unmodified reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py) and against excerpts of the
reference's shipped Voronoi masks.  Integer / bool / bin outputs must match bit for bit."""
import numpy as np
import pytest

import synth

F32 = np.float32


def bits(a):
    return np.ascontiguousarray(a, F32).view(np.uint32)


def assert_bitexact(a, b, what=""):
    a, b = np.asarray(a, F32), np.asarray(b, F32)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    same = (bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))
    assert same.all(), f"{what}: {np.count_nonzero(~same)} / {same.size} values differ bitwise"


# ----------------------------------------------------------------------------- stage 1
def test_ray_directions(orc, golden):
    g = golden("stage1")
    for cp in (True, False):
        d = orc.ray_directions(12, 16, 13.5, 14.25, 8.3, 5.9, cp)
        np.testing.assert_allclose(d, g[f"dirs_cp{int(cp)}"], atol=2e-7, rtol=0)


def test_aabb_intersect_bitexact(orc, golden):
    g = golden("stage1")
    o, d = synth.random_rays_in_box(11, 4096)
    tmin, tmax = orc.aabb_intersect(o, d, synth.AABB_GLOBAL)
    assert_bitexact(tmin, g["aabb_tmin"], "tmin")
    assert_bitexact(tmax, g["aabb_tmax"], "tmax")
    tmin, tmax = orc.aabb_intersect(o, d, synth.AABB_GLOBAL, invalid=float("inf"))
    assert_bitexact(tmin, g["aabb_tmin_inf"], "tmin_inf")
    assert_bitexact(tmax, g["aabb_tmax_inf"], "tmax_inf")
    assert np.isinf(tmin).any() and np.isfinite(tmin).any()   # both hit and miss cases covered


def test_get_rays_and_clamp(orc, golden):
    g = golden("stage1")
    cam = synth.nadir_rays(5, 1, H=24, W=32, f=25.0)[0]
    rays = orc.get_rays(g["cam_dirs"], cam["c2w"], aabb=synth.AABB_GLOBAL, invalid=float("inf"))
    ref = g["cam_rays"]
    np.testing.assert_allclose(rays[:, :6], ref[:, :6], atol=1e-6, rtol=0)
    # near/far are bit-exact once fed the reference's own o,d (the 3x3 rotation is float-tolerance)
    tmin, tmax = orc.aabb_intersect(ref[:, :3], ref[:, 3:6], synth.AABB_GLOBAL, invalid=float("inf"))
    assert_bitexact(tmin, ref[:, 6], "near")
    assert_bitexact(tmax, ref[:, 7], "far")
    for tag, ov in (("none", None), ("nn", (None, None)), ("nf", (0.05, 0.4)), ("n", (0.3, None))):
        r2, valid = orc.clamp_near_far(ref, ov)
        assert (valid == g[f"clamp_{tag}_valid"]).all(), tag
        assert_bitexact(r2, g[f"clamp_{tag}_rays"], f"clamp {tag}")
    rc = orc.get_rays(g["cam_dirs"], cam["c2w"], aabb=None, near=0.1, far=2.5)
    np.testing.assert_allclose(rc, g["rays_const"], atol=1e-6, rtol=0)


@pytest.mark.parametrize("S", [2, 3, 16, 17, 64, 65, 96, 255, 256])
def test_linspace_bitexact(orc, golden, S):
    assert_bitexact(orc.linspace01(S), golden("stage1")[f"linspace_{S}"], f"linspace {S}")


@pytest.mark.parametrize("S", [16, 64, 96])
def test_sample_bins_bitexact(orc, golden, S):
    g = golden("stage1")
    assert_bitexact(orc.stratified_t(g["t_rays"], S), g[f"t_eval_{S}"], "eval bins")
    assert_bitexact(orc.stratified_t(g["t_rays"], S, g[f"jitter_{S}"]), g[f"t_train_{S}"], "train bins")


def test_points_bitexact(orc, golden):
    g = golden("stage1")
    assert_bitexact(orc.points(g["t_rays"], g["t_train_64"]), g["pts_64"], "pts")


# ----------------------------------------------------------------------------- stage 2
def _hash_inputs():
    import importlib
    rng = np.random.default_rng(21)
    x = rng.uniform(0, 1, (2048, 3)).astype(F32)
    x[:8] = np.float32(1e-6)
    x[8:16] = np.float32(1.0) - np.float32(1e-6)
    x[16:24, 0] = np.float32(0.5)
    x[24:32] = (rng.integers(0, 16, (8, 3)) / 16.0).astype(F32)
    return x


def test_level_resolutions(orc, golden):
    g = golden("hashgrid")
    assert list(g["res"]) == [16, 23, 33, 48, 70, 101, 147, 212, 307, 445, 645, 933, 1351, 1955, 2830, 4095]
    for key in ("16_16_4096", "16_16_2048", "8_16_512", "4_16_4096", "1_16_4096"):
        L, mn, mx = (int(v) for v in key.split("_"))
        assert (orc.level_resolutions(L, mn, mx) == g[f"res_{key}"]).all(), key


def test_hash_indices_bitexact(orc, golden):
    g = golden("hashgrid")
    x = _hash_inputs()
    tab = np.zeros((16 << 12, 2), F32)
    _, idx = orc.hashgrid_fwd(x, tab, 16, 2, 12, g["res"], want_idx=True)
    assert (idx == g["idx_T12"]).all()
    for lt in (19, 20):
        tab = np.zeros((1, 2), F32)  # never dereferenced meaningfully: use a real-size dummy instead
        tab = np.zeros((16 << lt, 2), F32)
        _, idx = orc.hashgrid_fwd(x[:512], tab, 16, 2, lt, g["res"], want_idx=True)
        assert (idx == g[f"idx_T{lt}"]).all(), lt


@pytest.mark.parametrize("mode", ["Linear", "Smoothstep", "Nearest"])
def test_hashgrid_features_and_grad(orc, golden, mode):
    g = golden("hashgrid")
    x = _hash_inputs()
    sd = synth.make_expert_params(22, log2T=12)
    tab = sd["xyz_encoder.hash_table"]
    feat = orc.hashgrid_fwd(x, tab, 16, 2, 12, g["res"], mode)
    assert_bitexact(feat, g[f"feat_{mode}"], f"features {mode}")   # same op order -> same bits
    dout = np.random.default_rng(23).standard_normal((2048, 32)).astype(F32)
    dt = orc.hashgrid_bwd(x, dout, 16, 2, 12, g["res"], mode)
    np.testing.assert_allclose(dt, g[f"dtable_{mode}"], atol=2e-5, rtol=1e-5)


# ----------------------------------------------------------------------------- stage 3
def test_field_forward_and_backward(orc, golden):
    g = golden("field")
    sd = synth.make_expert_params(31, log2T=12)
    ws = synth.expert_weight_list(sd)
    lo, hi = synth.AABB_GLOBAL
    x01 = orc.world_to_unit(g["xyz"], lo, hi - lo)
    assert_bitexact(x01, g["x01"], "x01")
    res = orc.level_resolutions()
    enc = orc.hashgrid_fwd(x01, sd["xyz_encoder.hash_table"], 16, 2, 12, res)
    assert_bitexact(enc, g["enc"], "enc")
    np.testing.assert_allclose(orc.sh16(g["dirs"]), g["sh"], atol=2e-6, rtol=0)
    y = orc.field_fwd(enc, g["dirs"], ws)
    np.testing.assert_allclose(y[:, :3], g["y"][:, :3], atol=2e-6, rtol=0)
    np.testing.assert_allclose(y[:, 3], g["y"][:, 3], atol=0, rtol=2e-5)
    np.testing.assert_allclose(y[:, 3:4], g["sigma"], atol=0, rtol=2e-5)
    grads, d_enc = orc.field_bwd(enc, g["dirs"], ws, g["G"])
    for key, gr in zip(synth.EXPERT_KEYS, grads):
        ref = g["grad." + key]
        scale = np.abs(ref).max() + 1e-12
        assert np.abs(gr - ref).max() / scale < 2e-5, key
    dt = orc.hashgrid_bwd(x01, d_enc, 16, 2, 12, res)
    ref = g["grad.xyz_encoder.hash_table"]
    assert np.abs(dt - ref).max() / np.abs(ref).max() < 2e-5
    # the fp16-autocast emulation stays within the north-star RGB budget of the fp32 path
    yh = orc.field_fwd(enc, g["dirs"], ws, half=True)
    assert np.abs(yh[:, :3] - y[:, :3]).max() < 5e-3


# ----------------------------------------------------------------------------- stage 4
@pytest.mark.parametrize("tag,scale", [("bg", 1.0), ("nobg", 1.0), ("scale", 2.5)])
def test_composite(orc, golden, tag, scale):
    g = golden("composite")
    bg = None if tag == "nobg" else g["bg"]
    rgb, dep, w, acc = orc.composite_fwd(g["rgb_sigma"], g["t"], bg, scale)
    np.testing.assert_allclose(w, g[f"{tag}.weights"], atol=1e-6, rtol=1e-5)
    np.testing.assert_allclose(rgb, g[f"{tag}.rgb"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(dep, g[f"{tag}.depth"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(acc, g[f"{tag}.acc"], atol=2e-6, rtol=0)
    d, d_bg = orc.composite_bwd(g["rgb_sigma"], g["t"], bg, g["g_rgb"], g["g_depth"], g["g_weights"], g["g_acc"], scale)
    ref = g[f"{tag}.d_rgb_sigma"]
    np.testing.assert_allclose(d[..., :3], ref[..., :3], atol=2e-6, rtol=1e-5)
    err = np.abs(d[..., 3] - ref[..., 3])
    assert (err <= 1e-5 + 1e-4 * np.abs(ref[..., 3])).all(), err.max()
    if bg is not None:
        np.testing.assert_allclose(d_bg, g[f"{tag}.d_bg"], atol=2e-6, rtol=0)


# ----------------------------------------------------------------------------- stage 5
@pytest.mark.parametrize("tag,cen_key", [("g22", None), ("g24", "cen8")])
def test_point_routing_bitexact(orc, golden, tag, cen_key):
    g = golden("routing")
    cen = synth.CENTROIDS_G22 if cen_key is None else g[cen_key]
    _, hard = orc.route_points(g["pts"], cen, 1.0)
    assert (hard == g[f"{tag}.hard.1.0"]).all()
    for margin in (1.05, 1.1):
        w, _ = orc.route_points(g["pts"], cen, margin)
        ref = g[f"{tag}.w.{margin}"]
        assert ((w > 0) == (ref > 0)).all(), "support set"          # the integer/bool part: exact
        np.testing.assert_allclose(w, ref, atol=1.2e-7, rtol=0)     # FP weights: <= 2 ulp
        assert (bits(w) != bits(ref)).mean() < 1e-3


def test_point_routing_3d_tolerance(orc, golden):
    g = golden("routing")
    w, _ = orc.route_points(g["pts"], g["cen3d"], 1.05, cluster_2d=False)
    ref = g["g22_3d.w.1.05"]
    assert ((w > 0) != (ref > 0)).mean() < 1e-3
    agree = (w > 0) == (ref > 0)
    np.testing.assert_allclose(w[agree.all(1)], ref[agree.all(1)], atol=1e-5)


@pytest.mark.parametrize("stem", ["000005", "000007"])
def test_voronoi_vs_shipped_masks(orc, golden, stem):
    """Excerpts (~7k px per image incl. every kind of mask edge) of the masks the reference ships
    in data/drz/out/example/masks/g22_grid_bm110_ss11 -- the reference's own golden vectors."""
    g = golden("voronoi")
    rays = g[f"{stem}.rays"]
    mask = orc.route_rays_voronoi(rays, int(g["ray_samples"]), g["centroids"], float(g["margin"]))
    assert (mask == g[f"{stem}.voronoi_raw"]).all()
    _, valid = orc.clamp_near_far(rays, (None, None))
    assert (valid == g[f"{stem}.valid"]).all()
    assert ((mask & valid[:, None]) == g[f"{stem}.shipped"]).all()
    # stage 1 from the shipped camera metadata reproduces the rays the masks were cut from
    H, W = (int(v) for v in g[f"{stem}.HW"])
    fx, fy, cx, cy = (float(v) for v in g[f"{stem}.intrinsics"].astype(F32))
    dirs = orc.ray_directions(H, W, fx, fy, cx, cy, True).reshape(-1, 3)[g[f"{stem}.pix"]]
    mine = orc.get_rays(dirs, g[f"{stem}.c2w"], aabb=g["aabb"], invalid=float("inf"))
    fin = np.isfinite(rays[:, 6])
    np.testing.assert_allclose(mine[:, :6], rays[:, :6], atol=2e-6, rtol=0)
    np.testing.assert_allclose(mine[fin, 6:], rays[fin, 6:], atol=2e-4, rtol=1e-4)


def test_voronoi_other_margins(orc, golden):
    g = golden("voronoi")
    rays = g["000005.rays"][:2048]
    for margin in (1.0, 1.05):
        mask = orc.route_rays_voronoi(rays, 64, g["centroids"], margin)
        assert (mask == g[f"voronoi_m{margin}"]).all(), margin


# ----------------------------------------------------------------------------- glue
def _expert_setup(seed, log2T=12):
    sd = synth.make_expert_params(seed, log2T=log2T)
    lo, hi = synth.AABB_GLOBAL
    return sd, synth.expert_weight_list(sd), lo, hi - lo


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_render_expert_end_to_end(orc, golden, mode):
    g = golden("render")
    sd, ws, lo, ext = _expert_setup(83)
    res = orc.level_resolutions()
    jit = g["jitter"] if mode == "train" else None
    bg = np.ones((g["rays"].shape[0], 3), F32)     # bg_color_default="white"
    rgb, dep, w, acc, aux = orc.render_expert(g["rays"], 32, ws, sd["xyz_encoder.hash_table"], lo, ext,
                                              16, 2, 12, res, jitter=jit, bg=bg)
    np.testing.assert_allclose(rgb, g[f"{mode}.rgb"], atol=5e-6, rtol=0)
    np.testing.assert_allclose(dep, g[f"{mode}.depth"], atol=5e-6, rtol=0)
    np.testing.assert_allclose(w, g[f"{mode}.weights"], atol=5e-6, rtol=1e-4)
    np.testing.assert_allclose(acc, g[f"{mode}.acc"], atol=5e-6, rtol=0)
    # backward chain: composite -> field -> hash grid
    d_rs, _ = orc.composite_bwd(aux["rgb_sigma"].reshape(-1, 32, 4), aux["t"], bg, g["G_rgb"], g["G_depth"])
    grads, d_enc = orc.field_bwd(aux["enc"], aux["dirs"], ws, d_rs.reshape(-1, 4))
    for key, gr in zip(synth.EXPERT_KEYS, grads):
        ref = g[f"{mode}.grad.{key}"]
        assert np.abs(gr - ref).max() / (np.abs(ref).max() + 1e-12) < 1e-4, key
    dt = orc.hashgrid_bwd(aux["x01"], d_enc, 16, 2, 12, res)
    ref = g[f"{mode}.grad.table_sub"]
    assert np.abs(dt[::97] - ref).max() / np.abs(ref).max() < 1e-4
    dig = g[f"{mode}.grad.table_digest"]
    assert abs(np.abs(dt).sum(dtype=np.float64) - dig[1]) / dig[1] < 1e-4


@pytest.mark.parametrize("tag,margin", [("soft", 1.05), ("hard", 1.0)])
def test_container_blend(orc, golden, tag, margin):
    """4 experts, per-sample routing + blend + background MLP (meta_container.py:275-382)."""
    g = golden("render")
    rays, S = g["rays4"], 32
    res = orc.level_resolutions()
    t = orc.stratified_t(rays, S)
    pts = orc.points(rays, t).reshape(-1, 3)
    w, hard = orc.route_points(pts, synth.CENTROIDS_G22, margin)
    if w is not None:
        assert ((w > 0) == g["soft.support"]).all()
    else:
        assert (hard == g["hard.assign"]).all()
    dirs = np.repeat(rays[:, 3:6], S, axis=0)
    ys = []
    for k in range(4):
        sd = synth.make_expert_params(85 + k, log2T=12)
        lo, hi = synth.EXPERT_BOXES_G22[k]
        x01 = orc.world_to_unit(pts, lo, hi - lo)
        enc = orc.hashgrid_fwd(x01, sd["xyz_encoder.hash_table"], 16, 2, 12, res)
        ys.append(orc.field_fwd(enc, dirs, synth.expert_weight_list(sd)))
    y = orc.blend(np.stack(ys), w, hard)
    bgp = synth.make_bg_params(185)
    bg = orc.background(rays[:, 3:6], bgp["bg_mlp.0.weight"], bgp["bg_mlp.0.bias"], bgp["bg_mlp.2.weight"], bgp["bg_mlp.2.bias"])
    rgb, dep, wts, acc = orc.composite_fwd(y.reshape(-1, S, 4), t, bg)
    np.testing.assert_allclose(rgb, g[f"{tag}.rgb"], atol=5e-6, rtol=0)
    np.testing.assert_allclose(dep, g[f"{tag}.depth"], atol=5e-6, rtol=0)
    np.testing.assert_allclose(acc, g[f"{tag}.acc"], atol=5e-6, rtol=0)


# ----------------------------------------------------------------------------- loss epilogue / optimizer tail
@pytest.mark.parametrize("cs", ["linear", "srgb", "identity"])
def test_color_mse_vs_reference(orc, golden, cs):
    """color_space_transformer + F.mse_loss and autograd's gradient (nerfs/color_space.py, nerfs/losses.py:29-32)."""
    g = golden("loss")
    pred, gt = synth.loss_inputs()
    loss, dpred = orc.color_mse(pred, gt, cs, "mean")
    elem, _ = orc.color_mse(pred, gt, cs, "none")
    np.testing.assert_allclose(elem, g[f"{cs}_elem"], rtol=0, atol=5e-7)     # powf vs torch.pow differ by an ulp; squared
    np.testing.assert_allclose(loss, g[f"{cs}_loss"], rtol=2e-6)
    ref = g[f"{cs}_grad"]
    nan = np.isnan(ref)
    if cs == "srgb":
        # the reference's gradient is NaN exactly where pred == 0 (0 * inf in torch.where's unselected pow branch);
        # the restatement reports the selected linear branch there: 2 (0 - gt) * 12.92 / n
        assert (nan == (pred == 0.0)).all() and nan.sum() == 2
        want = 2.0 * (0.0 - np.clip(gt[nan], 0, 1)) * 12.92 / pred.size
        np.testing.assert_allclose(dpred[nan], want, rtol=1e-6)
    else:
        assert not nan.any()
    np.testing.assert_allclose(dpred[~nan], ref[~nan], rtol=2e-5, atol=1e-9)


@pytest.mark.parametrize("name,adamw,wd", [("adam", False, 0.0), ("adamw", True, 0.05), ("adam_wd", False, 0.05)])
def test_adam_step_vs_reference(orc, golden, name, adamw, wd):
    """get_optimizer + maml_meta_update (unscale_, clip_all_grads at 1.0, scaler.step, scaler.update) on CPU torch."""
    g = golden("optim")
    params, grads = synth.optim_inputs()
    lrs = [synth.OPTIM_LRS[grp] for grp, _ in synth.OPTIM_SHAPES]
    st = orc.AdamState(params, lrs, [wd] * len(params), adamw=adamw)
    scales = g[f"{name}_scales"]
    skipped = []
    for it, gs in enumerate(grads):
        with np.errstate(over="ignore"):
            scaled = [(a * F32(scales[it])).astype(F32) for a in gs]           # what scaler.scale(loss).backward() leaves in .grad
        norm, skip = st.update(scaled, grad_scale=float(scales[it]), max_norm=1.0)
        skipped.append(skip)
        if it in synth.OPTIM_KEEP:
            for k, p in enumerate(st.p):
                np.testing.assert_allclose(p, g[f"{name}_p{k}_step{it}"], rtol=2e-6, atol=2e-7, err_msg=f"step {it} tensor {k}")
    assert skipped == [False, False, False, True, False, False]
    assert st.step.value == float(g[f"{name}_steps"]) == 5.0
    if name == "adam":
        for k in range(len(st.p)):
            np.testing.assert_allclose(st.m[k], g[f"adam_m{k}"], rtol=2e-6, atol=1e-9)
            np.testing.assert_allclose(st.v[k], g[f"adam_v{k}"], rtol=2e-6, atol=1e-12)


# ----------------------------------------------------------------------------- task-grid binning (ray-batch producer)
def taskgrid_host_tensors(rays, cells, region=None):
    """The small host-side tensors of data/task_dataset.py: region box (:230-237 inferred from the near points when not
    given), clamped cell size (:241-245), per-cell bounds (:174-197) and tolerances (:595-597) -- torch ops, as there."""
    import torch
    r = torch.from_numpy(np.ascontiguousarray(rays, F32))
    if region is None:
        pts = r[:, 0:3] + r[:, 3:6] * r[:, 6:7]
        aabb = torch.stack([pts.min(dim=0).values, pts.max(dim=0).values], dim=0)
    else:
        aabb = torch.tensor(region, dtype=torch.float32)
    lo, hi = aabb[0], aabb[1]
    cell3 = torch.clamp((hi - lo) / torch.tensor(cells, dtype=torch.float32), min=1e-12)
    size = (hi - lo).clamp(min=1e-9)
    axes = [torch.linspace(0, 1, steps=n + 1) for n in cells]
    X0, Y0, Z0 = torch.meshgrid(*[a[:-1] for a in axes], indexing="ij")
    X1, Y1, Z1 = torch.meshgrid(*[a[1:] for a in axes], indexing="ij")
    lo_n = torch.stack([X0, Y0, Z0], dim=-1).reshape(-1, 3)
    hi_n = torch.stack([X1, Y1, Z1], dim=-1).reshape(-1, 3)
    bounds = torch.stack([lo + size * lo_n, lo + size * hi_n], dim=1)
    tol = torch.maximum(1e-6 * (bounds[:, 1] - bounds[:, 0]).norm(dim=1), torch.tensor(1e-9))
    return aabb.numpy(), cell3.numpy(), bounds.numpy(), tol.numpy()


@pytest.mark.parametrize("tag,cells,region", [("auto", (1, 6, 6), None), ("box", (2, 3, 4), "expert0")])
def test_dda_task_binning_vs_reference(orc, golden, tag, cells, region):
    """TaskDataset's "dda" routing (the policy nerf_runner.py:207 selects): the cell of every ray, bit for bit."""
    g = golden("taskgrid")
    rays = synth.task_rays()
    reg = None if region is None else tuple(map(tuple, synth.EXPERT_BOXES_G22[0].tolist()))
    aabb, cell3, bounds, tol = taskgrid_host_tensors(rays, cells, reg)
    assert_bitexact(aabb, g[f"{tag}_aabb"], "region box")
    assert_bitexact(bounds, g[f"{tag}_cell_bounds"], "cell bounds")
    cid, best_len, counts = orc.dda_route_rays(rays, aabb, cells, cell3, bounds, tol)
    assert (cid == g[f"{tag}_cid"]).all(), int((cid != g[f"{tag}_cid"]).sum())
    assert (counts == g[f"{tag}_counts"]).all()
    assert_bitexact(best_len, g[f"{tag}_best_len"], "in-cell length of the winning cell")
