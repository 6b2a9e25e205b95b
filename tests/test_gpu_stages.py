"""GPU parity tests, stage by stage: CUDA kernels (through the C ABI) vs the CPU oracle on the same
seeded inputs and vs the committed golden fixtures generated from the reference.

Bars (BASELINE north star): bit-exact for hash indices, sample bins, near/far, validity masks and
expert assignment; float tolerance (stated per test) for features, colours, depths, gradients."""
import numpy as np
import pytest
import torch

import synth
from helpers import F32, assert_bitexact, cu, npy, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from adaptive_city_nerf_b200 import ops as o
    return o


# ----------------------------------------------------------------------------- stage 1
def test_ray_directions(ops, golden):
    g = golden("stage1")
    from adaptive_city_nerf_b200.nerfs.ray_sampling import get_ray_directions
    for cp in (True, False):
        d = get_ray_directions(12, 16, 13.5, 14.25, 8.3, 5.9, cp, torch.device("cuda"))
        np.testing.assert_allclose(npy(d), g[f"dirs_cp{int(cp)}"], atol=2e-7, rtol=0)   # fp32 tolerance: 2 ulp


def test_aabb_intersect_bitexact(ops, golden, orc):
    from adaptive_city_nerf_b200.nerfs.scene_box import SceneBox
    g = golden("stage1")
    o, d = synth.random_rays_in_box(11, 4096)
    box = SceneBox(cu(synth.AABB_GLOBAL))
    tmin, tmax = box.ray_aabb_intersect(cu(o), cu(d))
    assert_bitexact(npy(tmin), g["aabb_tmin"], "tmin vs reference")
    assert_bitexact(npy(tmax), g["aabb_tmax"], "tmax vs reference")
    tmin, tmax = box.ray_aabb_intersect(cu(o), cu(d), invalid_value=float("inf"))
    assert_bitexact(npy(tmin), g["aabb_tmin_inf"], "tmin inf")
    # a different seed against the oracle
    o, d = synth.random_rays_in_box(12, 100_003)
    tmin, tmax = box.ray_aabb_intersect(cu(o), cu(d))
    omin, omax = orc.aabb_intersect(o, d, synth.AABB_GLOBAL)
    assert_bitexact(npy(tmin), omin, "tmin vs oracle")
    assert_bitexact(npy(tmax), omax, "tmax vs oracle")
    e0, e1 = box.ray_aabb_intersect(cu(o[:0]), cu(d[:0]))     # empty input
    assert e0.shape == (0,) and e1.shape == (0,)


def test_get_rays_and_clamp(ops, golden):
    from adaptive_city_nerf_b200.nerfs.ray_sampling import get_rays, clamp_rays_near_far
    from adaptive_city_nerf_b200.nerfs.scene_box import SceneBox
    g = golden("stage1")
    cam = synth.nadir_rays(5, 1, H=24, W=32, f=25.0)[0]
    box = SceneBox(cu(synth.AABB_GLOBAL))
    dirs = cu(g["cam_dirs"])
    rays = get_rays(dirs, cu(cam["c2w"]), scene_box=box, aabb_invalid_value=float("inf"))
    assert rays.shape == (24, 32, 8)
    rays = rays.view(-1, 8)
    ref = g["cam_rays"]
    np.testing.assert_allclose(npy(rays)[:, :6], ref[:, :6], atol=1e-6, rtol=0)
    # fed the reference's own o,d the clip is bit-exact
    tmin, tmax = box.ray_aabb_intersect(cu(ref[:, :3]), cu(ref[:, 3:6]), invalid_value=float("inf"))
    assert_bitexact(npy(tmin), ref[:, 6], "near")
    assert_bitexact(npy(tmax), ref[:, 7], "far")
    for tag, ov in (("none", None), ("nn", (None, None)), ("nf", (0.05, 0.4)), ("n", (0.3, None))):
        r2, valid = clamp_rays_near_far(cu(ref), ov)
        assert (npy(valid).astype(bool) == g[f"clamp_{tag}_valid"]).all(), tag
        assert_bitexact(npy(r2), g[f"clamp_{tag}_rays"], f"clamp {tag}")
    rc = get_rays(dirs.view(-1, 3), cu(cam["c2w"]), near=0.1, far=2.5)
    np.testing.assert_allclose(npy(rc), g["rays_const"], atol=1e-6, rtol=0)
    with pytest.raises(ValueError):
        get_rays(dirs.view(-1, 3), cu(cam["c2w"]))


@pytest.mark.parametrize("S", [16, 64, 96])
def test_sample_bins_bitexact(ops, golden, S):
    from adaptive_city_nerf_b200.nerfs.ray_rendering import stratified_t_vals
    g = golden("stage1")
    rays = cu(g["t_rays"])
    t = stratified_t_vals(rays[:, 6], rays[:, 7], S, randomized=False)
    assert_bitexact(npy(t), g[f"t_eval_{S}"], "eval bins")
    t = stratified_t_vals(rays[:, 6], rays[:, 7], S, randomized=True, jitter=cu(g[f"jitter_{S}"]))
    assert_bitexact(npy(t), g[f"t_train_{S}"], "train bins (shared jitter)")


def test_sample_bins_odd_S_vs_oracle(ops, orc, golden):
    rays = golden("stage1")["t_rays"]
    for S in (2, 3, 17, 65, 255):
        jit = np.random.default_rng(S).uniform(0, 1, (rays.shape[0], S)).astype(F32)
        assert_bitexact(npy(ops.sample_stratified(cu(rays), S, None)), orc.stratified_t(rays, S), f"S={S} eval")
        assert_bitexact(npy(ops.sample_stratified(cu(rays), S, cu(jit))), orc.stratified_t(rays, S, jit), f"S={S} train")


def test_points_bitexact(ops, golden):
    g = golden("stage1")
    id6 = ops.points(cu(g["t_rays"]), cu(g["t_train_64"]))
    assert_bitexact(npy(id6)[:, :3].reshape(-1, 64, 3), g["pts_64"], "pts")
    assert_bitexact(npy(id6)[:, 3:].reshape(-1, 64, 3)[:, 0], g["t_rays"][:, 3:6], "dirs")


# ----------------------------------------------------------------------------- stage 2
def _hash_inputs():
    rng = np.random.default_rng(21)
    x = rng.uniform(0, 1, (2048, 3)).astype(F32)
    x[:8] = np.float32(1e-6)
    x[8:16] = np.float32(1.0) - np.float32(1e-6)
    x[16:24, 0] = np.float32(0.5)
    x[24:32] = (rng.integers(0, 16, (8, 3)) / 16.0).astype(F32)
    return x


def _encoder(log2T, mode="Linear", table=None, L=16, F=2):
    from adaptive_city_nerf_b200.models.encodings import HashGridEncoder
    enc = HashGridEncoder(levels=L, min_res=16, max_res=4096, log2_hashmap_size=log2T, features_per_level=F,
                          interpolation=mode).cuda()
    if table is not None:
        with torch.no_grad():
            enc.hash_table.copy_(cu(table))
    return enc


@pytest.mark.parametrize("log2T", [12, 19, 20])
def test_hash_indices_bitexact(golden, log2T):
    g = golden("hashgrid")
    x = _hash_inputs()
    n = 2048 if log2T == 12 else 512
    enc = _encoder(log2T)
    idx = enc.hash_indices(cu(x[:n]))
    assert (npy(idx).astype(np.int64) == g[f"idx_T{log2T}"]).all()


@pytest.mark.parametrize("mode", ["Linear", "Smoothstep", "Nearest"])
def test_hashgrid_features_and_grad(golden, mode):
    g = golden("hashgrid")
    x = _hash_inputs()
    sd = synth.make_expert_params(22, log2T=12)
    enc = _encoder(12, mode, sd["xyz_encoder.hash_table"])
    y = enc(cu(x))
    assert y.shape == (2048, 32) and y.dtype == torch.float32
    assert_bitexact(npy(y), g[f"feat_{mode}"], f"features {mode}")     # same op order as the reference -> same bits
    dout = np.random.default_rng(23).standard_normal((2048, 32)).astype(F32)
    (y * cu(dout)).sum().backward()
    # atomics reorder the fp32 sums: tolerance 2e-5 abs on O(1)-magnitude sums of <= 2048 terms
    np.testing.assert_allclose(npy(enc.hash_table.grad), g[f"dtable_{mode}"], atol=2e-5, rtol=1e-5)


def test_hashgrid_shapes_and_edge_cases(orc):
    enc = _encoder(10)
    assert enc(torch.rand(0, 3, device="cuda")).shape == (0, 32)                # empty
    y = enc(torch.rand(3, 5, 3, device="cuda"))
    assert y.shape == (3, 5, 32)                                                # leading dims kept
    # features_per_level 1/4/8 and fewer levels against the oracle
    for F_, L_ in ((1, 4), (4, 8), (8, 2)):
        rng = np.random.default_rng(F_)
        tab = rng.uniform(-1, 1, (L_ << 10, F_)).astype(F32)
        e = _encoder(10, "Linear", tab, L=L_, F=F_)
        x = rng.uniform(0, 1, (777, 3)).astype(F32)
        res = npy(e.level_resolutions).astype(np.int32)
        assert_bitexact(npy(e(cu(x))), orc.hashgrid_fwd(x, tab, L_, F_, 10, res), f"F={F_}")
    # positions outside [0,1] (negative floors) hash like the reference's int64 arithmetic
    x = np.random.default_rng(9).uniform(-2, 3, (999, 3)).astype(F32)
    e = _encoder(12)
    res = npy(e.level_resolutions).astype(np.int32)
    _, idx = orc.hashgrid_fwd(x, np.zeros((16 << 12, 2), F32), 16, 2, 12, res, want_idx=True)
    assert (npy(e.hash_indices(cu(x))).astype(np.int64) == idx).all()


def test_hashgrid_full_size_properties():
    """BASELINE-size table (T=2^19, 64 MiB): linearity in the table and partition of unity."""
    enc = _encoder(19)
    x = torch.rand(1 << 18, 3, device="cuda")
    with torch.no_grad():
        enc.hash_table.fill_(1.0)
        y1 = enc(x)
        assert (y1 - 1.0).abs().max() < 2e-6          # trilinear weights sum to one
        enc.hash_table.copy_(torch.randn_like(enc.hash_table))
        a = enc(x)
        enc.hash_table.mul_(2.0)
        assert torch.equal(enc(x), 2.0 * a)           # exact linearity under power-of-two scaling
    idx = enc.hash_indices(x[:4096])
    lvl = torch.arange(16, device="cuda").view(1, 16, 1)
    assert ((idx >> 19) == lvl).all()                 # every level stays inside its own table slice


# ----------------------------------------------------------------------------- stage 3
def _field_inputs(golden):
    g = golden("field")
    sd = synth.make_expert_params(31, log2T=12)
    return g, sd, synth.expert_weight_list(sd)


def test_sh16(ops, golden):
    g = golden("field")
    np.testing.assert_allclose(npy(ops.sh16(cu(g["dirs"]))), g["sh"], atol=2e-6, rtol=0)


def test_field_fp32_forward_backward(ops, golden, orc):
    g, sd, ws = _field_inputs(golden)
    wt = [cu(w).requires_grad_(True) for w in ws]
    enc = cu(g["enc"])
    dirs = cu(g["dirs"])
    y = ops.field_fwd(enc, dirs, 3, 1, wt, half=False)
    np.testing.assert_allclose(npy(y)[:, :3], g["y"][:, :3], atol=1e-5, rtol=0)     # fp32 path: 1e-5 abs on rgb
    np.testing.assert_allclose(npy(y)[:, 3], g["y"][:, 3], atol=0, rtol=1e-4)       # sigma spans decades: relative
    grads, d_enc = ops.field_bwd(enc, dirs, 3, 1, wt, False, cu(g["G"]), True, [True] * 14)
    for key, gr in zip(synth.EXPERT_KEYS, grads):
        assert rel_err(npy(gr), g["grad." + key]) < 1e-4, key
    ogr, o_denc = orc.field_bwd(g["enc"], g["dirs"], ws, g["G"])
    assert rel_err(npy(d_enc), o_denc) < 1e-4
    # ragged sizes: tails of the 64-point tiles
    for P in (1, 63, 65, 200):
        y2 = ops.field_fwd(enc[:P].contiguous(), dirs[:P].contiguous(), 3, 1, wt, half=False)
        np.testing.assert_allclose(npy(y2), npy(y)[:P], atol=1e-6, rtol=1e-6)


def test_expert_forward_autograd_fp32(golden):
    """MetaNGP.forward through autograd: all 14 MLP tensors + the hash table (reference grads)."""
    from helpers import make_container
    g = golden("field")
    m = make_container(1, np.zeros((1, 3)), [synth.AABB_GLOBAL], 1.0, False, 31)
    ex = m.submodules[0]
    x6 = cu(np.concatenate([g["xyz"], g["dirs"]], axis=1))
    y = ex(x6)
    np.testing.assert_allclose(npy(y)[:, :3], g["y"][:, :3], atol=1e-5, rtol=0)
    (y * cu(g["G"])).sum().backward()
    named = dict(ex.named_parameters())
    for key in synth.EXPERT_KEYS:
        assert rel_err(npy(named[key].grad), g["grad." + key]) < 1e-4, key
    assert rel_err(npy(ex.xyz_encoder.hash_table.grad), g["grad.xyz_encoder.hash_table"]) < 1e-4
    # split API agrees with the fused kernels
    with torch.no_grad():
        dens = ex.density(cu(g["xyz"]), return_feats=True)
        np.testing.assert_allclose(npy(dens["sigma"]), g["sigma"], rtol=1e-4, atol=0)
        np.testing.assert_allclose(npy(dens["geo_feat"]), g["geo"], atol=2e-5, rtol=1e-4)
        rgb = ex.color(cu(g["dirs"]), dens["geo_feat"])
        np.testing.assert_allclose(npy(rgb), g["y"][:, :3], atol=1e-5, rtol=0)


# ----------------------------------------------------------------------------- stage 4
@pytest.mark.parametrize("tag,scale", [("bg", 1.0), ("nobg", 1.0), ("scale", 2.5)])
def test_composite(golden, tag, scale):
    from adaptive_city_nerf_b200.nerfs.ray_rendering import volume_render
    g = golden("composite")
    rs = cu(g["rgb_sigma"]).requires_grad_(True)
    bg = None if tag == "nobg" else cu(g["bg"]).requires_grad_(True)
    rgb, dep, w, acc = volume_render(rs, cu(g["t"]), bg_rgb=bg, sigma_scale=scale)
    # fp32 compositing; scan order differs from the reference's sequential cumprod: 2e-6 abs
    np.testing.assert_allclose(npy(w), g[f"{tag}.weights"], atol=2e-6, rtol=1e-5)
    np.testing.assert_allclose(npy(rgb), g[f"{tag}.rgb"], atol=3e-6, rtol=0)
    np.testing.assert_allclose(npy(dep), g[f"{tag}.depth"], atol=3e-6, rtol=0)
    np.testing.assert_allclose(npy(acc), g[f"{tag}.acc"], atol=3e-6, rtol=0)
    loss = (rgb * cu(g["g_rgb"])).sum() + (dep * cu(g["g_depth"])).sum() + (w * cu(g["g_weights"])).sum() + (acc * cu(g["g_acc"])).sum()
    loss.backward()
    ref = g[f"{tag}.d_rgb_sigma"]
    d = npy(rs.grad)
    np.testing.assert_allclose(d[..., :3], ref[..., :3], atol=3e-6, rtol=1e-5)
    err = np.abs(d[..., 3] - ref[..., 3])
    assert (err <= 2e-5 + 2e-4 * np.abs(ref[..., 3])).all(), err.max()
    if bg is not None:
        np.testing.assert_allclose(npy(bg.grad), g[f"{tag}.d_bg"], atol=3e-6, rtol=0)


def test_composite_long_rays_and_properties(orc):
    """S not a multiple of 32, S = 256, early termination, and weights summing to acc."""
    from adaptive_city_nerf_b200.nerfs.ray_rendering import volume_render
    rng = np.random.default_rng(5)
    for S in (2, 31, 33, 100, 256):
        N = 257
        rs = rng.uniform(0, 1, (N, S, 4)).astype(F32)
        rs[..., 3] = rng.uniform(0, 400 if S > 64 else 20, (N, S))          # dense -> saturates -> early exit
        t = np.sort(rng.uniform(0, 1, (N, S)), axis=1).astype(F32)
        bg = rng.uniform(0, 1, (N, 3)).astype(F32)
        rgb, dep, w, acc = volume_render(cu(rs), cu(t), bg_rgb=cu(bg))
        o = orc.composite_fwd(rs, t, bg)
        np.testing.assert_allclose(npy(rgb), o[0], atol=3e-6, rtol=0)
        np.testing.assert_allclose(npy(w), o[2], atol=2e-6, rtol=1e-5)
        np.testing.assert_allclose(npy(w).sum(1), npy(acc), atol=2e-6)
        assert (npy(acc) <= 1.0 + 1e-6).all()
    z = volume_render(torch.rand(0, 8, 4, device="cuda"), torch.rand(0, 8, device="cuda"))
    assert z[0].shape == (0, 3) and z[2].shape == (0, 8)


# ----------------------------------------------------------------------------- stage 5
@pytest.mark.parametrize("tag", ["g22", "g24"])
def test_point_routing(ops, golden, tag):
    g = golden("routing")
    cen = synth.CENTROIDS_G22 if tag == "g22" else g["cen8"]
    pts = cu(g["pts"])
    _, hard, counts = ops.route_points(pts, cu(cen), 2, 1.0, want_counts=True)
    assert (npy(hard).astype(np.int64) == g[f"{tag}.hard.1.0"]).all()                  # bit-exact assignment
    assert (npy(counts) == np.bincount(g[f"{tag}.hard.1.0"], minlength=cen.shape[0])).all()
    for margin in (1.05, 1.1):
        w, _, counts = ops.route_points(pts, cu(cen), 2, margin, want_counts=True)
        ref = g[f"{tag}.w.{margin}"]
        assert ((npy(w) > 0) == (ref > 0)).all(), "support set"                        # bit-exact support
        np.testing.assert_allclose(npy(w), ref, atol=1.2e-7, rtol=0)                   # FP weights: <= 2 ulp
        assert (npy(counts) == (ref > 0).sum(0)).all()


def test_point_routing_vs_oracle_large(ops, orc):
    rng = np.random.default_rng(77)
    lo, hi = synth.AABB_GLOBAL
    pts = (lo + rng.uniform(0, 1, (300_001, 3)) * (hi - lo)).astype(F32)
    for margin in (1.0, 1.05):
        w, h, _ = ops.route_points(cu(pts), cu(synth.CENTROIDS_G22), 2, margin)
        ow, oh = orc.route_points(pts, synth.CENTROIDS_G22, margin)
        if margin == 1.0:
            assert (npy(h) == oh).all()
        else:
            assert_bitexact(npy(w), ow, "soft weights vs oracle")      # same op order as the C oracle
    # 3-D clustering (tolerance-only in the reference too)
    w, _, _ = ops.route_points(cu(pts[:5000]), cu(synth.CENTROIDS_G22), 3, 1.05)
    ow, _ = orc.route_points(pts[:5000], synth.CENTROIDS_G22, 1.05, cluster_2d=False)
    assert_bitexact(npy(w), ow, "3-D weights vs oracle")


@pytest.mark.parametrize("stem", ["000005", "000007"])
def test_voronoi_masks_vs_shipped(ops, golden, stem):
    """The reference's own golden vectors: excerpts of data/drz/out/example/masks/g22_grid_bm110_ss11."""
    from adaptive_city_nerf_b200.nerfs.ray_sampling import clamp_rays_near_far
    g = golden("voronoi")
    rays = cu(g[f"{stem}.rays"])
    mask = ops.route_rays_voronoi(rays, int(g["ray_samples"]), cu(g["centroids"]), 2, float(g["margin"]))
    assert (npy(mask).astype(bool) == g[f"{stem}.voronoi_raw"]).all()
    _, valid = clamp_rays_near_far(rays, (None, None))
    assert ((npy(mask).astype(bool) & npy(valid).astype(bool)[:, None]) == g[f"{stem}.shipped"]).all()


def test_voronoi_other_margins_and_k8(ops, golden, orc):
    g = golden("voronoi")
    rays = g["000005.rays"][:2048]
    for margin in (1.0, 1.05):
        mask = ops.route_rays_voronoi(cu(rays), 64, cu(g["centroids"]), 2, margin)
        assert (npy(mask).astype(bool) == g[f"voronoi_m{margin}"]).all(), margin
    cen8 = golden("routing")["cen8"]
    fin = np.isfinite(rays[:, 6])
    mask = ops.route_rays_voronoi(cu(rays), 256, cu(cen8), 2, 1.05)
    assert (npy(mask).astype(bool)[fin] == orc.route_rays_voronoi(rays, 256, cen8, 1.05)[fin]).all()


def test_bucket_and_blend(ops, orc):
    rng = np.random.default_rng(3)
    lo, hi = synth.AABB_GLOBAL
    P = 10_007
    id6 = np.concatenate([(lo + rng.uniform(0, 1, (P, 3)) * (hi - lo)), rng.standard_normal((P, 3))], 1).astype(F32)
    for margin in (1.0, 1.1):
        w, h, counts = ops.route_points(cu(id6), cu(synth.CENTROIDS_G22), 2, margin, want_counts=True)
        cnt = npy(counts).astype(np.int64)
        off = np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.int32)
        sel, xd, ws = ops.bucket_points(cu(id6), w, h, 4, cu(off), int(cnt.sum()))
        sel_n, xd_n, ws_n = npy(sel).astype(np.int64), npy(xd), npy(ws)
        wn = npy(w) if w is not None else np.eye(4, dtype=F32)[npy(h).astype(np.int64)]
        for k in range(4):
            s = sel_n[off[k]:off[k] + cnt[k]]
            assert sorted(s.tolist()) == np.nonzero(wn[:, k] > 0)[0].tolist()       # same set as nonzero()
            assert (xd_n[off[k]:off[k] + cnt[k]] == id6[s]).all()
            assert (ws_n[off[k]:off[k] + cnt[k]] == wn[s, k]).all()
        y = rng.standard_normal((int(cnt.sum()), 4)).astype(F32)
        out = torch.zeros(P, 4, device="cuda")
        for k in range(4):
            sl = slice(int(off[k]), int(off[k] + cnt[k]))
            out = ops.BlendFn.apply(out, cu(y[sl]), ws[sl], sel[sl])
        ref = np.zeros((P, 4), F32)
        for k in range(4):
            sl = slice(int(off[k]), int(off[k] + cnt[k]))
            np.add.at(ref, sel_n[sl], y[sl] * ws_n[sl, None])
        np.testing.assert_allclose(npy(out), ref, atol=1e-6)


@pytest.mark.parametrize("K,margin,dims", [(4, 1.05, 2), (4, 1.0, 2), (8, 1.1, 2), (3, 1.05, 3), (13, 1.05, 2)])
def test_route_and_bucket_straight_from_rays(ops, orc, golden, K, margin, dims):
    """acn_route_count_rays / acn_route_bucket_rays == acn_points + acn_route_points + acn_bucket_points, bit for bit
    (rows, weights, per-expert sets), and the counts equal the oracle's routing of the oracle's points."""
    rng = np.random.default_rng(40 + K)
    N, S = 1531, 24                                                      # N*S not a multiple of the block size
    o, d = synth.random_rays_in_box(50 + K, N)
    rays_np = np.concatenate([o, d, np.zeros((N, 1), F32), rng.uniform(0.2, 0.6, (N, 1)).astype(F32)], 1).astype(F32)
    rays = cu(rays_np)
    t = ops.sample_stratified(rays, S, cu(rng.uniform(0, 1, (N, S)).astype(F32)))
    cen_np = synth.CENTROIDS_G22[:K] if K <= 4 else np.concatenate(
        [np.zeros((K, 1)), rng.uniform(-1, 1, (K, 2))], 1).astype(F32)
    cen = cu(cen_np)
    id6 = ops.points(rays, t)
    w, h, counts = ops.route_points(id6, cen, dims, margin, want_counts=True)
    cnt_f = ops.route_count_rays(rays, t, cen, dims, margin)
    assert (npy(cnt_f) == npy(counts)).all()
    ow, oh = orc.route_points(npy(id6)[:, :3], cen_np, margin, cluster_2d=dims == 2)
    o_cnt = (ow > 0).sum(0) if margin > 1 else np.bincount(oh, minlength=K)
    assert (npy(cnt_f).astype(np.int64) == o_cnt).all()                  # bit-exact support sets vs the CPU oracle
    cnt = npy(counts).astype(np.int64)
    off = np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.int32)
    sel_a, xd_a, w_a = ops.bucket_points(id6, w, h, K, cu(off), int(cnt.sum()))
    cnt_s, support = ops.route_count_rays(rays, t, cen, dims, margin, want_support=True)
    assert (npy(cnt_s) == npy(counts)).all()
    wn = npy(w) > 0 if w is not None else np.eye(K, dtype=bool)[npy(h).astype(np.int64)]
    assert (support.cpu().numpy().astype(np.int64) == (wn * (1 << np.arange(K))).sum(1)).all()     # one bit per expert in the set
    for sup in (None, support):                                          # bucket pass: routing recomputed / read back
        for ray_major in (False, True):                                  # row order inside a bucket; the sets are the same
            sel_b, xd_b, w_b = ops.route_bucket_rays(rays, t, cen, dims, margin, cu(off), int(cnt.sum()), support=sup,
                                                     ray_major=ray_major)
            _compare_buckets(K, off, cnt, (sel_a, xd_a, w_a), (sel_b, xd_b, w_b))
    cnt_r, support_r = ops.route_count_rays(rays, t, cen, dims, margin, want_support=True, ray_major=True)
    assert (npy(cnt_r) == npy(counts)).all() and torch.equal(support_r, support)
    empty = ops.route_count_rays(rays[:0], t[:0], cen, dims, margin)
    assert int(empty.sum()) == 0


def test_route_from_rays_on_the_margin_boundary(ops, orc):
    """Samples placed ON the surfaces d_k = margin * d_min (Apollonius circles of centroid pairs, rounded to fp32 so they
    fall on either side by an ulp): the support sets from the fused kernels (which skip the square roots of clearly distant
    experts) must equal the oracle's evaluation of every distance, bit for bit."""
    rng = np.random.default_rng(12)
    cen = synth.CENTROIDS_G22.astype(np.float64)
    for margin in (1.05, 1.1, 1.5):
        pts = []
        for i in range(4):
            for j in range(4):
                if i == j:
                    continue
                c0, c1 = cen[i, 1:], cen[j, 1:]
                m2 = np.float64(np.float32(margin)) ** 2
                ctr, rad = (m2 * c0 - c1) / (m2 - 1.0), np.sqrt(m2) * np.linalg.norm(c0 - c1) / (m2 - 1.0)
                th = rng.uniform(0, 2 * np.pi, 60000)
                q = ctr + rad * np.stack([np.cos(th), np.sin(th)], 1)
                q = q[(np.abs(q) < 1.2).all(1)]
                pts.append(q)
        yz = np.concatenate(pts)
        yz = np.concatenate([yz, yz * (1 + rng.normal(0, 3e-7, yz.shape))])      # and a few ulps around them
        N = yz.shape[0]
        o = np.concatenate([rng.uniform(0, 0.4, (N, 1)), yz], 1).astype(F32)
        rays_np = np.concatenate([o, np.zeros((N, 3), F32), np.zeros((N, 1), F32), np.ones((N, 1), F32)], 1).astype(F32)
        rays = cu(rays_np)
        t = torch.full((N, 1), 0.5, device="cuda")
        counts, support = ops.route_count_rays(rays, t, cu(synth.CENTROIDS_G22), 2, margin, want_support=True)
        ow, _ = orc.route_points(o, synth.CENTROIDS_G22, margin)
        want = ((ow > 0) * (1 << np.arange(4))).sum(1)
        got = support.cpu().numpy().astype(np.int64)
        assert (got == want).all(), (margin, int((got != want).sum()))
        assert ((ow > 0).sum(1) >= 2).mean() > 0.2            # the boundary really is exercised
        assert (npy(counts) == (ow > 0).sum(0)).all()


def _compare_buckets(K, off, cnt, a, b):
    (sel_a, xd_a, w_a), (sel_b, xd_b, w_b) = a, b
    for k in range(K):
        sl = slice(int(off[k]), int(off[k] + cnt[k]))
        oa, ob = np.argsort(npy(sel_a[sl]).astype(np.int64)), np.argsort(npy(sel_b[sl]).astype(np.int64))
        assert (npy(sel_a[sl])[oa] == npy(sel_b[sl])[ob]).all()
        assert_bitexact(npy(xd_a[sl])[oa], npy(xd_b[sl])[ob], f"rows of expert {k}")
        assert_bitexact(npy(w_a[sl])[oa], npy(w_b[sl])[ob], f"weights of expert {k}")


def test_hashgrid_fwd_rays_ray_major_is_the_same_encoding(ops):
    """ray_major only changes which thread encodes which sample (frames: a warp = one sample of 32 adjacent pixels)."""
    from adaptive_city_nerf_b200.nerfs.ray_sampling import get_rays, get_ray_directions
    from adaptive_city_nerf_b200.nerfs.scene_box import SceneBox
    box = SceneBox(cu(synth.AABB_GLOBAL))
    cam = synth.nadir_rays(5, 1, H=37, W=53, f=400.0)[0]                   # N = 1961 rays: not a multiple of 32
    dirs = get_ray_directions(37, 53, cam["fx"], cam["fy"], cam["cx"], cam["cy"], True, torch.device("cuda"))
    rays = get_rays(dirs, cu(cam["c2w"]), scene_box=box).view(-1, 8).contiguous()
    spec = ops.GridSpec(16, 2, 14, _encoder(14).level_resolutions.clone(), 1)
    table = (torch.rand(16 << 14, 2, device="cuda", generator=torch.Generator(device="cuda").manual_seed(4)) - 0.5)
    box6 = torch.cat([box.min, box.extent]).contiguous()
    for S in (13, 64):                                                     # S not a multiple of 8 / a multiple
        t = ops.sample_stratified(rays, S, None)
        for dt in (torch.float16, torch.float32):
            a = ops.hashgrid_fwd_rays(rays, t, table, spec, box6, dt, ray_major=False)
            b = ops.hashgrid_fwd_rays(rays, t, table, spec, box6, dt, ray_major=True)
            assert torch.equal(a, b)


def test_hashgrid_bwd_march_vs_generic(ops):
    """The ray-marching scatter (run-length aggregation, rotated start) must equal the generic
    per-(point,level) scatter: same sums, different order -> fp32 tolerance."""
    from adaptive_city_nerf_b200.nerfs.ray_sampling import get_rays, get_ray_directions
    from adaptive_city_nerf_b200.nerfs.scene_box import SceneBox
    box = SceneBox(cu(synth.AABB_GLOBAL))
    cam = synth.nadir_rays(17, 1, H=48, W=33, f=25.0)[0]
    dirs = get_ray_directions(48, 33, cam["fx"], cam["fy"], cam["cx"], cam["cy"], True, torch.device("cuda"))
    rays = get_rays(dirs, cu(cam["c2w"]), scene_box=box).view(-1, 8)
    for S, log2T, mode in ((64, 14, 1), (37, 12, 2), (2, 10, 1)):
        N = rays.shape[0]
        t = ops.sample_stratified(rays, S, torch.rand(N, S, device="cuda"))
        enc = _encoder(log2T)
        spec = ops.GridSpec(16, 2, log2T, enc.level_resolutions.clone(), mode)
        box6 = torch.cat([box.min, box.extent]).contiguous()
        gen = torch.Generator(device="cuda").manual_seed(S)
        dout = torch.randn(N * S, 32, device="cuda", generator=gen)
        dout[::7] = 0.0                                     # zero rows are skipped
        a = torch.zeros(16 << log2T, 2, device="cuda")
        b = torch.zeros_like(a)
        ops.hashgrid_bwd_rays(rays, t, dout, spec, box6, a)                       # march kernel (rays)
        ops.hashgrid_bwd_plain(ops.points(rays, t), dout, spec, box6, b)          # plain per-(point, level) kernel
        scale = float(b.abs().max())
        assert float((a - b).abs().max()) <= 2e-5 * scale + 1e-6, (S, log2T, mode)
        # march kernel over a point list (the routed path): in order, shuffled (every row its own cell), fp16, short tail
        pts = ops.points(rays, t)
        for order in (None, torch.randperm(N * S, device="cuda", generator=gen)):
            pp, dd = (pts, dout) if order is None else (pts[order].contiguous(), dout[order].contiguous())
            c = torch.zeros_like(a)
            ops.hashgrid_bwd(pp, dd, spec, box6, c)
            assert float((c - b).abs().max()) <= 2e-5 * scale + 1e-6, (S, log2T, mode, order is None)
        c16 = torch.zeros_like(a)
        ops.hashgrid_bwd(pts[:-37], dout[:-37].half(), spec, box6, c16)
        d_ref = torch.zeros_like(a)
        ops.hashgrid_bwd_plain(pts[:-37], dout[:-37], spec, box6, d_ref)
        assert float((c16 - d_ref).abs().max()) <= 2e-3 * scale
        # fp16 dL/denc input
        a16 = torch.zeros_like(a)
        ops.hashgrid_bwd_rays(rays, t, dout.half(), spec, box6, a16)
        assert float((a16 - b).abs().max()) <= 2e-3 * scale
