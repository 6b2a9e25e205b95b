"""Generate the committed golden fixtures by running the UNMODIFIED reference in-process.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
Two import shims (nerfacc, viser.transforms) are needed and touch nothing on the hot path
(SURVEY Appendix A).  tinycudann is absent, so every encoder takes the reference's pure-torch
branch -- exactly the CPU path that BASELINE config 1 names.  Inputs come from tests/golden/synth.py
(numpy PCG64) so the tests can regenerate them; only OUTPUTS are stored.
"""
import importlib.util
import io
import sys
import types
import warnings
import zipfile
from collections import OrderedDict
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import synth  # noqa: E402

REF = Path("/root/reference")
na = types.ModuleType("nerfacc"); na.OccGridEstimator = object
sys.modules["nerfacc"] = na
v = types.ModuleType("viser"); vt = types.ModuleType("viser.transforms"); v.transforms = vt
sys.modules["viser"] = v; sys.modules["viser.transforms"] = vt
sys.path.insert(0, str(REF))
warnings.filterwarnings("ignore")

from nerfs.scene_box import SceneBox  # noqa: E402
from nerfs.ray_sampling import get_ray_directions, get_rays, clamp_rays_near_far  # noqa: E402
from nerfs.ray_rendering import render_rays, stratified_t_vals, volume_render  # noqa: E402
from models.encodings import HashGridEncoder, SHEncoder  # noqa: E402
from models.inr.meta_container import MetaContainer  # noqa: E402

torch.set_num_threads(8)
T = torch.from_numpy
F32 = np.float32


def save(name, **arrs):
    np.savez_compressed(HERE / name, **{k: np.asarray(v) for k, v in arrs.items()})
    sz = (HERE / name).stat().st_size
    print(f"  {name}: {sz / 1024:.1f} KiB, keys={list(arrs)[:12]}{' ...' if len(arrs) > 12 else ''}")


# --------------------------------------------------------------------------- stage 1
def gen_stage1():
    out = {}
    H, W, fx, fy, cx, cy = 12, 16, 13.5, 14.25, 8.3, 5.9
    for cp in (True, False):
        out[f"dirs_cp{int(cp)}"] = get_ray_directions(H, W, fx, fy, cx, cy, cp, torch.device("cpu")).numpy()
    box = SceneBox(aabb=T(synth.AABB_GLOBAL))
    o, d = synth.random_rays_in_box(11, 4096)
    tmin, tmax = box.ray_aabb_intersect(T(o), T(d))
    out["aabb_tmin"], out["aabb_tmax"] = tmin.numpy(), tmax.numpy()
    tmin, tmax = box.ray_aabb_intersect(T(o), T(d), invalid_value=float("inf"))
    out["aabb_tmin_inf"], out["aabb_tmax_inf"] = tmin.numpy(), tmax.numpy()
    # get_rays through a synthetic nadir camera
    cam = synth.nadir_rays(5, 1, H=24, W=32, f=25.0)[0]
    dirs = get_ray_directions(cam["H"], cam["W"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], True, torch.device("cpu"))
    rays = get_rays(dirs, T(cam["c2w"]), scene_box=box, aabb_invalid_value=float("inf")).view(-1, 8)
    out["cam_rays"] = rays.numpy()
    out["cam_dirs"] = dirs.numpy()
    for tag, ov in (("none", None), ("nn", (None, None)), ("nf", (0.05, 0.4)), ("n", (0.3, None))):
        r2, valid = clamp_rays_near_far(rays, ov)
        out[f"clamp_{tag}_rays"], out[f"clamp_{tag}_valid"] = r2.numpy(), valid.numpy()
    # rays with constant near/far (no box)
    out["rays_const"] = get_rays(dirs.view(-1, 3), T(cam["c2w"]), near=0.1, far=2.5).numpy()
    for S in (2, 3, 16, 17, 64, 65, 96, 255, 256):
        out[f"linspace_{S}"] = torch.linspace(0.0, 1.0, S).numpy()
    # stratified t: eval + train with captured jitter
    r2, valid = clamp_rays_near_far(rays, (None, None))
    rv = r2[valid][:300]
    for S in (16, 64, 96):
        out[f"t_eval_{S}"] = stratified_t_vals(rv[:, 6], rv[:, 7], S, randomized=False).numpy()
        torch.manual_seed(100 + S)
        out[f"t_train_{S}"] = stratified_t_vals(rv[:, 6], rv[:, 7], S, randomized=True).numpy()
        torch.manual_seed(100 + S)
        out[f"jitter_{S}"] = torch.rand(rv.shape[0], S).numpy()
    out["t_rays"] = rv.numpy()
    tv = T(out["t_train_64"])
    pts = rv[:, None, :3] + rv[:, None, 3:6] * tv[..., None]
    out["pts_64"] = pts.numpy()
    save("stage1.npz", **out)


# --------------------------------------------------------------------------- stage 2
def hash_inputs(seed, P):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 1, (P, 3)).astype(F32)
    x[:8] = np.float32(1e-6)
    x[8:16] = np.float32(1.0) - np.float32(1e-6)
    x[16:24, 0] = np.float32(0.5)      # exact cell boundaries on even resolutions
    x[24:32] = (rng.integers(0, 16, (8, 3)) / 16.0).astype(F32)
    return x


def ref_indices(enc, x01):
    L = enc.levels
    levels = enc.level_resolutions.to(dtype=x01.dtype)
    scaled = x01[..., None, :] * levels.view(1, L, 1)
    fl = torch.floor(scaled).to(torch.int64)
    ce = fl + 1
    idx = []
    for cxb in (0, 1):
        for cyb in (0, 1):
            for czb in (0, 1):
                ix = (ce if cxb else fl)[..., 0]
                iy = (ce if cyb else fl)[..., 1]
                iz = (ce if czb else fl)[..., 2]
                idx.append(enc._hash(ix, iy, iz) + enc.level_offsets)
    return torch.stack(idx, dim=-1)  # (P,L,8) order 000,001,...,111 (x,y,z bits)


def gen_hashgrid():
    out = {}
    P, L, F, log2T = 2048, 16, 2, 12
    x = hash_inputs(21, P)
    sd = synth.make_expert_params(22, L=L, F=F, log2T=log2T)
    rng = np.random.default_rng(23)
    dout = rng.standard_normal((P, L * F)).astype(F32)
    for mode in ("Linear", "Smoothstep", "Nearest"):
        enc = HashGridEncoder(levels=L, min_res=16, max_res=4096, log2_hashmap_size=log2T,
                              features_per_level=F, interpolation=mode)
        assert not enc._use_tcnn
        with torch.no_grad():
            enc.hash_table.copy_(T(sd["xyz_encoder.hash_table"]))
        y = enc(T(x))
        (y * T(dout)).sum().backward()
        out[f"feat_{mode}"] = y.detach().numpy()
        out[f"dtable_{mode}"] = enc.hash_table.grad.numpy()
        if mode == "Linear":
            out["res"] = enc.level_resolutions.numpy()
            out["idx_T12"] = ref_indices(enc, T(x)).numpy().astype(np.int32)
    for lt in (19, 20):
        enc = HashGridEncoder(levels=1, log2_hashmap_size=lt)  # tiny table; only the hash matters
        enc.levels = 16
        big = HashGridEncoder.__new__(HashGridEncoder)
        torch.nn.Module.__init__(big)
        big.levels, big.log2_hashmap_size = 16, lt
        big.register_buffer("level_resolutions", T(out["res"]))
        big.register_buffer("level_offsets", torch.arange(16, dtype=torch.int64) * (2 ** lt))
        big.register_buffer("hash_primes", torch.tensor([1, 2654435761, 805459861], dtype=torch.int64))
        out[f"idx_T{lt}"] = ref_indices(big, T(x[:512])).numpy().astype(np.int32)
    # other level configurations (resolutions only)
    for (Lc, mn, mx) in ((16, 16, 4096), (16, 16, 2048), (8, 16, 512), (4, 16, 4096), (1, 16, 4096)):
        e = HashGridEncoder(levels=Lc, min_res=mn, max_res=mx, log2_hashmap_size=4)
        out[f"res_{Lc}_{mn}_{mx}"] = e.level_resolutions.numpy()
    save("hashgrid.npz", **out)


# --------------------------------------------------------------------------- experts / container
HASH_CONF = dict(levels=16, features_per_level=2, log2_hashmap_size=12, max_res=4096, min_res=16,
                 interpolation="Linear")


def make_container(K, centroids, boxes, margin, use_bg, seed0, hash_conf=HASH_CONF, cluster_2d=True):
    m = MetaContainer(
        num_submodules=K, centroids=T(np.asarray(centroids, F32)), aabb=T(synth.AABB_GLOBAL),
        boundary_margin=margin, cluster_2d=cluster_2d, use_bg_nerf=use_bg,
        expert_box_list=[SceneBox(aabb=T(np.asarray(b, F32))) for b in boxes],
        hidden=64, sigma_depth=2, color_depth=2, color_hidden=64, dir_encoding="spherical",
        use_sigmoid_rgb=True, hash_enc_conf=dict(hash_conf), occ_conf={"use_occ": False})
    sd = m.state_dict()
    for k in range(K):
        p = synth.make_expert_params(seed0 + k, L=hash_conf["levels"], F=hash_conf["features_per_level"],
                                     log2T=hash_conf["log2_hashmap_size"])
        for key, val in p.items():
            sd[f"submodules.{k}.{key}"] = T(val)
    if use_bg:
        for key, val in synth.make_bg_params(seed0 + 100).items():
            sd[key] = T(val)
    m.load_state_dict(sd)
    return m


def expert_rays(seed, n):
    cams = synth.nadir_rays(seed, 1, H=16, W=n // 16, f=14.0)
    cam = cams[0]
    box = SceneBox(aabb=T(synth.AABB_GLOBAL))
    dirs = get_ray_directions(cam["H"], cam["W"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], True, torch.device("cpu"))
    rays = get_rays(dirs, T(cam["c2w"]), scene_box=box).view(-1, 8)
    rays, valid = clamp_rays_near_far(rays, (None, None))
    assert bool(valid.all())
    return rays


def gen_field():
    """MetaNGP.forward on world points + grads of a random linear functional."""
    out = {}
    m = make_container(1, np.zeros((1, 3)), [synth.AABB_GLOBAL], 1.0, False, 31)
    ex = m.submodules[0]
    rng = np.random.default_rng(32)
    P = 1024
    lo, hi = synth.AABB_GLOBAL
    xyz = (lo + rng.uniform(-0.02, 1.02, (P, 3)) * (hi - lo)).astype(F32)  # a few outside -> clamp
    d = rng.standard_normal((P, 3)).astype(F32)
    d[:4] *= 1e-3
    G = rng.standard_normal((P, 4)).astype(F32)
    x6 = T(np.concatenate([xyz, d], axis=1))
    y = ex(x6)
    (y * T(G)).sum().backward()
    out["xyz"], out["dirs"], out["G"], out["y"] = xyz, d, G, y.detach().numpy()
    for key in synth.EXPERT_KEYS:
        out["grad." + key] = dict(ex.named_parameters())[key].grad.numpy()
    out["grad.xyz_encoder.hash_table"] = ex.xyz_encoder.hash_table.grad.numpy()
    # intermediate pieces for stage-wise checks
    with torch.no_grad():
        x01 = ex._world_to_unit(T(xyz))
        out["x01"] = x01.numpy()
        out["enc"] = ex.xyz_encoder(x01).numpy()
        out["sh"] = ex._enc_dir(T(d)).numpy()
        dens = ex.density(T(xyz), return_feats=True)
        out["sigma"], out["geo"] = dens["sigma"].numpy(), dens["geo_feat"].numpy()
    save("field.npz", **out)


def gen_composite():
    out = {}
    rng = np.random.default_rng(41)
    N, S = 256, 48
    rs = rng.uniform(-0.2, 1.2, (N, S, 4)).astype(F32)
    rs[..., 3] = (rng.standard_normal((N, S)) * 30).astype(F32)   # mixed sign, large
    rs[:32, :, 3] = np.abs(rs[:32, :, 3]) * 20                    # opaque early -> tests saturation
    rs[32:48, :, 3] = -1.0                                         # empty rays
    near = rng.uniform(0, 0.1, (N, 1)); far = near + rng.uniform(0.2, 0.6, (N, 1))
    t = np.sort(near + (far - near) * rng.uniform(0, 1, (N, S)), axis=1).astype(F32)
    t[48:64, 5] = t[48:64, 4]                                      # zero-length intervals -> 1e-4 clamp
    bg = rng.uniform(0, 1, (N, 3)).astype(F32)
    gs = [rng.standard_normal(s).astype(F32) for s in ((N, 3), (N,), (N, S), (N,))]
    for tag, b, scale in (("bg", bg, 1.0), ("nobg", None, 1.0), ("scale", bg, 2.5)):
        rst = T(rs).requires_grad_(True)
        bgt = T(b).requires_grad_(True) if b is not None else None
        o = volume_render(rst, T(t), bg_rgb=bgt, sigma_scale=scale)
        loss = sum((oi * T(g)).sum() for oi, g in zip(o, gs))
        loss.backward()
        for name, oi in zip(("rgb", "depth", "weights", "acc"), o):
            out[f"{tag}.{name}"] = oi.detach().numpy()
        out[f"{tag}.d_rgb_sigma"] = rst.grad.numpy()
        if bgt is not None:
            out[f"{tag}.d_bg"] = bgt.grad.numpy()
    out.update(rgb_sigma=rs, t=t, bg=bg, g_rgb=gs[0], g_depth=gs[1], g_weights=gs[2], g_acc=gs[3])
    save("composite.npz", **out)


def grid_centroids_2x4():
    cams = np.array([[-0.04, -0.9, -0.9], [-0.04, 0.9, 0.9]], F32)
    spec = importlib.util.spec_from_file_location("cc", REF / "scripts" / "create_clusters.py")
    return spec, cams


def load_cc():
    for name in ("data.dataset", "data.image_metadata"):
        pass
    spec = importlib.util.spec_from_file_location("create_clusters_ref", REF / "scripts" / "create_clusters.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def gen_routing(cc):
    out = {}
    rng = np.random.default_rng(51)
    P = 8192
    lo, hi = synth.AABB_GLOBAL
    pts = (lo + rng.uniform(0, 1, (P, 3)) * (hi - lo)).astype(F32)
    pts[:16] = synth.CENTROIDS_G22[rng.integers(0, 4, 16)]            # exactly on a centroid
    pts[16:64, 1] = 0.0                                               # on the Voronoi edge
    pts[64:128, 2] = 0.0
    cen8 = cc._grid_centroids(T(np.array([[-0.04, -0.9, -0.9], [-0.04, 0.9, 0.9]], F32)), 1, 2, 4, True).numpy().astype(F32)
    out["pts"], out["cen8"] = pts, cen8
    for tag, cen in (("g22", synth.CENTROIDS_G22), ("g24", cen8)):
        K = cen.shape[0]
        for margin in (1.0, 1.05, 1.1):
            m = make_container(K, cen, [synth.AABB_GLOBAL] * K, margin, False, 61,
                               hash_conf=dict(HASH_CONF, levels=1, log2_hashmap_size=4))
            w, h = m._routing(T(pts))
            if w is not None:
                out[f"{tag}.w.{margin}"] = w.numpy()
            else:
                out[f"{tag}.hard.{margin}"] = h.numpy().astype(np.int32)
    m = make_container(4, synth.CENTROIDS_G22 + np.array([[0.1, 0, 0], [0.3, 0, 0], [0.2, 0, 0], [0.0, 0, 0]], F32),
                       [synth.AABB_GLOBAL] * 4, 1.05, False, 61,
                       hash_conf=dict(HASH_CONF, levels=1, log2_hashmap_size=4), cluster_2d=False)
    out["cen3d"] = m.centroids.numpy()
    out["g22_3d.w.1.05"] = m._routing(T(pts))[0].numpy()
    save("routing.npz", **out)


def read_zip_mask(path):
    with zipfile.ZipFile(path, "r") as zf:
        name = zf.namelist()[0]
        with zf.open(name) as f:
            return torch.load(io.BytesIO(f.read()), map_location="cpu")


def gen_voronoi(cc):
    """Excerpts of the reference's SHIPPED masks (g22_grid_bm110_ss11) + compute_voronoi_orig."""
    out = {}
    root = REF / "data/drz/out/example"
    mdir = root / "masks/g22_grid_bm110_ss11"
    params = torch.load(mdir / "params.pt")
    cents = params["centroids"]
    out["centroids"] = cents.numpy()
    out["margin"] = np.float32(params["boundary_margin"])
    out["ray_samples"] = np.int32(params["ray_samples"])
    box = SceneBox(aabb=params["aabb_global"])
    out["aabb"] = params["aabb_global"].numpy()
    rng = np.random.default_rng(71)
    for stem in ("000005", "000007"):
        md = torch.load(root / "train/metadata" / f"{stem}.pt", map_location="cpu")
        H, W = int(md["H"]), int(md["W"])
        fx, fy, cx, cy = md["intrinsics"]
        dirs = get_ray_directions(H, W, fx, fy, cx, cy, True, torch.device("cpu"))
        rays = get_rays(dirs, md["c2w"], scene_box=box, aabb_max_bound=1e10,
                        aabb_invalid_value=float("inf")).view(-1, 8)
        rays, valid = clamp_rays_near_far(rays, (None, None))
        shipped = torch.stack([read_zip_mask(mdir / str(c) / f"{stem}.pt").view(-1) for c in range(4)], dim=1)
        # pick pixels: random + every pixel whose membership differs from its right neighbour (edges)
        sm = shipped.view(H, W, 4)
        edge = (sm[:, 1:] != sm[:, :-1]).any(-1)
        edge_idx = torch.nonzero(edge.reshape(-1)).squeeze(1)
        edge_lin = (edge_idx // (W - 1)) * W + (edge_idx % (W - 1))
        pick_e = edge_lin[T(rng.choice(len(edge_lin), size=min(3000, len(edge_lin)), replace=False))] if len(edge_lin) else edge_lin
        pick_r = T(rng.choice(H * W, size=5000, replace=False))
        pick = torch.unique(torch.cat([pick_e, pick_r]))
        sub = rays[pick]
        vor = cc.compute_voronoi_orig(sub, ray_samples=256, ray_chunk_size=32768, sample_chunk_size=2 ** 29,
                                      centroids=cents, cluster_2d=True, device=torch.device("cpu"),
                                      boundary_margin=float(params["boundary_margin"]))
        mine = vor & valid[pick].unsqueeze(1)
        mism = int((mine != shipped[pick]).sum())
        print(f"  voronoi {stem}: {len(pick)} px, reference-vs-shipped mismatches = {mism}")
        assert mism == 0
        out[f"{stem}.pix"] = pick.numpy().astype(np.int32)
        out[f"{stem}.rays"] = sub.numpy()
        out[f"{stem}.valid"] = valid[pick].numpy()
        out[f"{stem}.shipped"] = shipped[pick].numpy()
        out[f"{stem}.voronoi_raw"] = vor.numpy()
        out[f"{stem}.c2w"] = md["c2w"].numpy()
        out[f"{stem}.intrinsics"] = np.array([float(a) for a in md["intrinsics"]], np.float64)
        out[f"{stem}.HW"] = np.array([H, W], np.int32)
    # margin 1.05 and hard rule have no shipped masks: run the reference
    sub = T(out["000005.rays"][:2048])
    for margin in (1.0, 1.05):
        out[f"voronoi_m{margin}"] = cc.compute_voronoi_orig(
            sub, ray_samples=64, ray_chunk_size=32768, sample_chunk_size=2 ** 29, centroids=cents,
            cluster_2d=True, device=torch.device("cpu"), boundary_margin=margin).numpy()
    save("voronoi.npz", **out)


def grads_of(model, loss):
    model.zero_grad(set_to_none=True)
    loss.backward()
    return {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}


def table_digest(g):
    """Per-level sums / abs-sums plus a fixed strided subsample of a (L*T,F) table gradient."""
    g = g.numpy()
    return g[:: 97].copy(), np.array([g.sum(dtype=np.float64), np.abs(g).sum(dtype=np.float64)])


def gen_render():
    """render_rays end to end: one expert (active_module=0) eval + train, fast weights, and a
    4-expert soft-routed container with the background MLP."""
    out = {}
    rays = expert_rays(81, 256)
    out["rays"] = rays.numpy()
    S = 32
    rngG = np.random.default_rng(82)
    Gr = rngG.standard_normal((rays.shape[0], 3)).astype(F32)
    Gd = rngG.standard_normal((rays.shape[0],)).astype(F32)
    m = make_container(1, np.zeros((1, 3)), [synth.AABB_GLOBAL], 1.0, False, 83)
    for mode in ("eval", "train"):
        m.train(mode == "train")
        torch.manual_seed(7)
        o = render_rays(m, rays, ray_samples=S, active_module=0, chunk=1 << 20)
        if mode == "train":
            torch.manual_seed(7)
            out["jitter"] = torch.rand(rays.shape[0], S).numpy()
        loss = (o[0] * T(Gr)).sum() + (o[1] * T(Gd)).sum()
        g = grads_of(m, loss)
        for name, oi in zip(("rgb", "depth", "weights", "acc"), o):
            out[f"{mode}.{name}"] = oi.detach().numpy()
        for key in synth.EXPERT_KEYS:
            out[f"{mode}.grad.{key}"] = g[f"submodules.0.{key}"].numpy()
        sub, dig = table_digest(g["submodules.0.xyz_encoder.hash_table"])
        out[f"{mode}.grad.table_sub"], out[f"{mode}.grad.table_digest"] = sub, dig
    out["G_rgb"], out["G_depth"] = Gr, Gd
    # fast weights (params=): with active_module set render_rays hands `params` straight to the
    # expert (ray_rendering.py:323-325), so keys are expert-relative (meta_core.py:27,196-205)
    m.eval()
    fast = OrderedDict((n, (p * 1.25).detach().requires_grad_(True))
                       for n, p in m.submodules[0].meta_named_parameters())
    o = render_rays(m, rays, ray_samples=S, params=fast, active_module=0)
    out["fast.rgb"] = o[0].detach().numpy()
    gr = torch.autograd.grad((o[0] * T(Gr)).sum(), list(fast.values()))
    for (n, _), gi in zip(fast.items(), gr):
        out["fast.grad." + n] = gi.numpy()
    out["fast.keys"] = np.array(list(fast.keys()))
    # 4-expert container, soft routing 1.05 and hard routing, background MLP
    box = SceneBox(aabb=T(synth.AABB_GLOBAL))
    cam = synth.nadir_rays(84, 1, H=16, W=16, f=7.0)[0]     # wide FOV: crosses all four cells
    cam["c2w"] = synth.camera_c2w(0.03, -0.02, 0.3)
    dirs = get_ray_directions(16, 16, cam["fx"], cam["fy"], cam["cx"], cam["cy"], True, torch.device("cpu"))
    rays4 = get_rays(dirs, T(cam["c2w"]), scene_box=box).view(-1, 8)
    rays4, valid = clamp_rays_near_far(rays4, (None, None))
    rays4 = rays4[valid]
    out["rays4"] = rays4.numpy()
    G4 = rngG.standard_normal((rays4.shape[0], 3)).astype(F32)
    out["G4"] = G4
    for tag, margin in (("soft", 1.05), ("hard", 1.0)):
        mc = make_container(4, synth.CENTROIDS_G22, synth.EXPERT_BOXES_G22, margin, True, 85)
        mc.eval()
        o = render_rays(mc, rays4, ray_samples=S, active_module=None, chunk=1 << 20)
        g = grads_of(mc, (o[0] * T(G4)).sum())
        for name, oi in zip(("rgb", "depth", "weights", "acc"), o):
            out[f"{tag}.{name}"] = oi.detach().numpy()
        for k in range(4):
            for key in ("sigma_trunk.0.linear.weight", "color_mlp.2.bias", "sigma_head.weight"):
                out[f"{tag}.grad.{k}.{key}"] = g[f"submodules.{k}.{key}"].numpy()
            sub, dig = table_digest(g[f"submodules.{k}.xyz_encoder.hash_table"])
            out[f"{tag}.grad.{k}.table_sub"], out[f"{tag}.grad.{k}.table_digest"] = sub, dig
        for key in ("bg_mlp.0.weight", "bg_mlp.2.bias"):
            out[f"{tag}.grad.{key}"] = g[key].numpy()
        with torch.no_grad():
            pts = (rays4[:, None, :3] + rays4[:, None, 3:6] * stratified_t_vals(rays4[:, 6], rays4[:, 7], S, False)[..., None]).reshape(-1, 3)
            w, h = mc._routing(pts)
            if w is not None:
                out[f"{tag}.support"] = (w > 0).numpy()
            else:
                out[f"{tag}.assign"] = h.numpy().astype(np.int32)
    save("render.npz", **out)


def gen_train():
    """Fixed-step Adam training of one expert on an analytic scene; PSNR trajectory."""
    out = {}
    steps, N, S = 150, 1024, 32
    conf = dict(HASH_CONF, log2_hashmap_size=14)
    m = make_container(1, np.zeros((1, 3)), [synth.AABB_GLOBAL], 1.0, False, 91, hash_conf=conf)
    # reference init scale for the table (encodings.py:267) so the optimisation is realistic
    with torch.no_grad():
        m.submodules[0].xyz_encoder.hash_table.mul_(1e-3 / 0.5)
    groups = m.get_param_groups()
    opt = torch.optim.Adam([
        {"params": groups["encoding"]["params"], "lr": 1e-2},
        {"params": groups["sigma"]["params"], "lr": 2e-3},
        {"params": groups["color"]["params"], "lr": 2e-3}], eps=1e-15)
    cams = synth.nadir_rays(92, 8, H=32, W=32, f=30.0)
    box = SceneBox(aabb=T(synth.AABB_GLOBAL))
    all_rays = []
    for cam in cams:
        dirs = get_ray_directions(cam["H"], cam["W"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], True, torch.device("cpu"))
        r = get_rays(dirs, T(cam["c2w"]), scene_box=box).view(-1, 8)
        r, valid = clamp_rays_near_far(r, (None, None))
        all_rays.append(r[valid])
    all_rays = torch.cat(all_rays)
    # analytic target: colour of the ground point hit at x = 0.45 (a smooth function of y,z)
    tg = (0.45 - all_rays[:, 0]) / all_rays[:, 3]
    hit = all_rays[:, :3] + all_rays[:, 3:6] * tg[:, None]
    gt = torch.stack([0.5 + 0.5 * torch.sin(3 * hit[:, 1]), 0.5 + 0.5 * torch.cos(2 * hit[:, 2]),
                      0.5 + 0.25 * torch.sin(2 * hit[:, 1] + hit[:, 2])], dim=1).clamp(0, 1)
    out["all_rays"], out["gt"] = all_rays.numpy(), gt.numpy()
    rng = np.random.default_rng(93)
    batch_idx = rng.integers(0, all_rays.shape[0], (steps, N))
    out["batch_idx"] = batch_idx.astype(np.int32)
    m.train()
    torch.manual_seed(1234)
    jit = torch.rand(steps, N, S)
    out["jitter_seed"] = np.int64(1234)
    psnr = []
    import nerfs.ray_rendering as rr
    for it in range(steps):
        idx = T(batch_idx[it])
        rays, tgt = all_rays[idx], gt[idx]
        # feed the captured jitter: stratified_t_vals draws rand_like(low) -> patch torch.rand_like once
        orig = torch.rand_like
        torch.rand_like = lambda x, _j=jit[it]: _j
        try:
            rgb, *_ = render_rays(m, rays, ray_samples=S, active_module=0)
        finally:
            torch.rand_like = orig
        loss = torch.nn.functional.mse_loss(rgb, tgt)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        psnr.append(-10.0 * np.log10(float(loss) + 1e-24))
    out["psnr"] = np.array(psnr, np.float64)
    m.eval()
    with torch.no_grad():
        rgb, *_ = render_rays(m, all_rays[:2048], ray_samples=S, active_module=0)
        out["final_eval_psnr"] = np.float64(-10.0 * np.log10(float(torch.nn.functional.mse_loss(rgb, gt[:2048])) + 1e-24))
    print("  train psnr first/last:", psnr[0], psnr[-1], "eval:", out["final_eval_psnr"])
    save("train.npz", **out)


# --------------------------------------------------------------------------- loss epilogue / optimizer tail
def gen_loss():
    """nerfs/color_space.py color_space_transformer + F.mse_loss (nerfs/losses.py:29-32) and autograd's d loss/d pred."""
    from nerfs.color_space import color_space_transformer
    pred_np, gt_np = synth.loss_inputs()
    out = {}
    for cs in ("linear", "srgb", "identity"):
        p = T(pred_np).clone().requires_grad_()
        a, b = color_space_transformer(p, T(gt_np), cs)
        loss = torch.nn.functional.mse_loss(a, b, reduction="mean")
        loss.backward()
        out[f"{cs}_loss"] = loss.detach().numpy()
        out[f"{cs}_grad"] = p.grad.numpy()          # sRGB: NaN where pred == 0 (0 * inf in the unselected pow branch)
        out[f"{cs}_elem"] = torch.nn.functional.mse_loss(a, b, reduction="none").detach().numpy()
    save("loss.npz", **out)


def gen_optim():
    """common/utils.py get_optimizer + pipelines/offline_stage/meta_core.py maml_meta_update (GradScaler.unscale_,
    clip_all_grads, scaler.step, scaler.update) on CPU, for Adam and AdamW."""
    from common.utils import get_optimizer
    from pipelines.offline_stage.meta_core import maml_meta_update
    import contextlib, io as _io
    params_np, grads_np = synth.optim_inputs()
    out = {}
    for name, wd in (("adam", 0.0), ("adamw", 0.05), ("adam_wd", 0.05)):
        class Holder(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.ps = torch.nn.ParameterList([torch.nn.Parameter(T(a).clone()) for a in params_np])
            def get_param_groups(self):
                g = {}
                for (grp, _), p in zip(synth.OPTIM_SHAPES, self.ps):
                    g.setdefault(grp, {"params": []})["params"].append(p)
                return g
        m = Holder()
        P = types.SimpleNamespace(lr=1e-3, encoding_lr=synth.OPTIM_LRS["encoding"], sigma_lr=synth.OPTIM_LRS["sigma"],
                                  color_lr=synth.OPTIM_LRS["color"], bg_lr=synth.OPTIM_LRS["background"],
                                  optimizer="adamw" if name == "adamw" else "adam", weight_decay=wd)
        opt = get_optimizer(P, m)
        scaler = torch.amp.GradScaler("cpu", init_scale=65536.0, growth_interval=2)
        scales = []
        for it, gs in enumerate(grads_np):
            scales.append(scaler.get_scale())
            loss = sum((p * T(g)).sum() for p, g in zip(m.ps, gs))
            with contextlib.redirect_stdout(_io.StringIO()):
                maml_meta_update(opt, loss, scaler, grad_clip=1.0)
            if it in synth.OPTIM_KEEP:
                for k, p in enumerate(m.ps):
                    out[f"{name}_p{k}_step{it}"] = p.detach().numpy().copy()
        out[f"{name}_scales"] = np.array(scales + [scaler.get_scale()], np.float64)
        if name == "adam":
            for k, p in enumerate(m.ps):
                out[f"{name}_m{k}"] = opt.state[p]["exp_avg"].numpy().copy()
                out[f"{name}_v{k}"] = opt.state[p]["exp_avg_sq"].numpy().copy()
        out[f"{name}_steps"] = np.float64(float(opt.state[m.ps[0]]["step"]))
    save("optim.npz", **out)


def gen_taskgrid():
    """data/task_dataset.py TaskDataset "dda" routing (nerf_runner.py:201-209): per-ray task cell, the bins, the region
    box inferred from the near points, and the host-side tensors (cell bounds, tolerances) the binning uses."""
    from data.task_dataset import TaskDataset
    rays = T(synth.task_rays())
    ram = types.SimpleNamespace(_rays=rays, _rgbs=torch.zeros(rays.shape[0], 3), _img_indices=torch.arange(rays.shape[0]) // 1600)
    out = {}
    for tag, cells, region in (("auto", (1, 6, 6), None),
                               ("box", (2, 3, 4), tuple(map(tuple, synth.EXPERT_BOXES_G22[0].tolist())))):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ds = TaskDataset(ram_ds=ram, cell_id=0, S_target=64, Q_target=32, min_rays_cell=10, image_cap=0.4,
                             assignment_checkpoint=0.7, routing_policy="dda", cells=cells, region_bounds=region)
        bins = ds._route_and_bin(ds.rays, ds.aabb, ds.cells, 0.7)
        cid = np.full(rays.shape[0], -1, np.int32)
        for c, b in enumerate(bins):
            cid[b.numpy()] = c
        valid, t0, t1, seg = ds._region_segment(ds.rays, ds.aabb)
        best = np.zeros(rays.shape[0], F32)
        _, bl = ds._dda_maxoverlap(ds.rays[valid], t0[valid], t1[valid], max_steps=64)
        best[valid.numpy()] = bl.numpy()
        out[f"{tag}_cid"], out[f"{tag}_best_len"], out[f"{tag}_valid"] = cid, best, valid.numpy()
        out[f"{tag}_aabb"], out[f"{tag}_cell_bounds"] = ds.aabb.numpy(), ds.cell_bounds.numpy()
        out[f"{tag}_counts"] = np.array([b.numel() for b in bins], np.int64)
        print(f"  taskgrid[{tag}]: {int(valid.sum())} valid of {rays.shape[0]}, {int((cid >= 0).sum())} binned, "
              f"cells used {int((out[f'{tag}_counts'] > 0).sum())}/{len(bins)}")
    save("taskgrid.npz", **out)


# --------------------------------------------------------------------------- round 2: the benchmarked arithmetic and configs 3-5
def train_scene():
    """The rays / targets / batches / jitter of gen_train (same seeds), regenerated instead of stored twice."""
    steps, N, S = 150, 1024, 32
    cams = synth.nadir_rays(92, 8, H=32, W=32, f=30.0)
    box = SceneBox(aabb=T(synth.AABB_GLOBAL))
    all_rays = []
    for cam in cams:
        dirs = get_ray_directions(cam["H"], cam["W"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], True, torch.device("cpu"))
        r = get_rays(dirs, T(cam["c2w"]), scene_box=box).view(-1, 8)
        r, valid = clamp_rays_near_far(r, (None, None))
        all_rays.append(r[valid])
    all_rays = torch.cat(all_rays)
    tg = (0.45 - all_rays[:, 0]) / all_rays[:, 3]
    hit = all_rays[:, :3] + all_rays[:, 3:6] * tg[:, None]
    gt = torch.stack([0.5 + 0.5 * torch.sin(3 * hit[:, 1]), 0.5 + 0.5 * torch.cos(2 * hit[:, 2]),
                      0.5 + 0.25 * torch.sin(2 * hit[:, 1] + hit[:, 2])], dim=1).clamp(0, 1)
    batch_idx = np.random.default_rng(93).integers(0, all_rays.shape[0], (steps, N))
    torch.manual_seed(1234)
    jit = torch.rand(steps, N, S)
    return steps, N, S, all_rays, gt, batch_idx, jit


def with_jitter(j, fn):
    orig = torch.rand_like
    torch.rand_like = lambda x, _j=j: _j
    try:
        return fn()
    finally:
        torch.rand_like = orig


def gen_train_amp():
    """The reference's mixed-precision training step -- torch.autocast(float16) around the loss, GradScaler, global clip,
    Adam: pipelines/offline_stage/meta_core.py:123-141 maml_meta_update (called unmodified) -- on the gen_train scene,
    150 steps.  On a GPU the reference enters torch.cuda.amp.autocast; here the same autocast policy runs on the CPU
    (matmul -> fp16 with fp32 accumulation, everything else as listed), which is the arithmetic the tcgen05 path must
    track: PSNR trajectory + final eval PSNR."""
    from pipelines.offline_stage.meta_core import maml_meta_update
    import contextlib, io as _io
    out = {}
    steps, N, S, all_rays, gt, batch_idx, jit = train_scene()
    conf = dict(HASH_CONF, log2_hashmap_size=14)
    m = make_container(1, np.zeros((1, 3)), [synth.AABB_GLOBAL], 1.0, False, 91, hash_conf=conf)
    with torch.no_grad():
        m.submodules[0].xyz_encoder.hash_table.mul_(1e-3 / 0.5)
    groups = m.get_param_groups()
    opt = torch.optim.Adam([
        {"params": groups["encoding"]["params"], "lr": 1e-2},
        {"params": groups["sigma"]["params"], "lr": 2e-3},
        {"params": groups["color"]["params"], "lr": 2e-3}], eps=1e-15)
    scaler = torch.amp.GradScaler("cpu", init_scale=65536.0)
    m.train()
    psnr, scales = [], []
    for it in range(steps):
        idx = T(batch_idx[it])
        rays, tgt = all_rays[idx], gt[idx]
        scales.append(scaler.get_scale())
        with torch.autocast("cpu", dtype=torch.float16):
            rgb, *_ = with_jitter(jit[it], lambda: render_rays(m, rays, ray_samples=S, active_module=0))
            loss = torch.nn.functional.mse_loss(rgb, tgt)
        with contextlib.redirect_stdout(_io.StringIO()):
            maml_meta_update(opt, loss, scaler, grad_clip=1.0)
        psnr.append(-10.0 * np.log10(float(loss) + 1e-24))
    out["psnr"] = np.array(psnr, np.float64)
    out["scales"] = np.array(scales + [scaler.get_scale()], np.float64)
    m.eval()
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.float16):
        rgb, *_ = render_rays(m, all_rays[:2048], ray_samples=S, active_module=0)
        out["final_eval_psnr"] = np.float64(-10.0 * np.log10(float(torch.nn.functional.mse_loss(rgb.float(), gt[:2048])) + 1e-24))
    print("  train_amp psnr first/last:", psnr[0], psnr[-1], "eval:", out["final_eval_psnr"], "scale:", scales[0], "->", scales[-1])
    save("train_amp.npz", **out)


def rays_over_scene(seed, n_views, H, W, f, yaw_spread=True):
    box = SceneBox(aabb=T(synth.AABB_GLOBAL))
    all_rays = []
    for cam in synth.nadir_rays(seed, n_views, H=H, W=W, f=f):
        dirs = get_ray_directions(cam["H"], cam["W"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], True, torch.device("cpu"))
        r = get_rays(dirs, T(cam["c2w"]), scene_box=box).view(-1, 8)
        r, valid = clamp_rays_near_far(r, (None, None))
        all_rays.append(r[valid])
    return torch.cat(all_rays)


def gen_render8():
    """BASELINE config 4 in small: the 2x4 grid (8 experts, margin 1.05, background head) rendered through
    render_rays(..., active_module=None) -- models/inr/meta_container.py:275-343 -- in eval mode, with gradients of a
    random functional for every expert and the background head."""
    out = {}
    cen = synth.CENTROIDS_G24
    boxes = synth.expert_boxes_for_grid(cen)
    S = 32
    # three wide-angle views whose footprints cover all eight cells
    rays = rays_over_scene(201, 3, 20, 20, 6.0)
    out["rays"] = rays.numpy()
    rng = np.random.default_rng(202)
    G = rng.standard_normal((rays.shape[0], 3)).astype(F32)
    Gd = rng.standard_normal((rays.shape[0],)).astype(F32)
    out["G_rgb"], out["G_depth"] = G, Gd
    mc = make_container(8, cen, boxes, 1.05, True, 203)
    mc.eval()
    o = render_rays(mc, rays, ray_samples=S, active_module=None, chunk=1 << 20)
    g = grads_of(mc, (o[0] * T(G)).sum() + (o[1] * T(Gd)).sum())
    for name, oi in zip(("rgb", "depth", "weights", "acc"), o):
        out[name] = oi.detach().numpy()
    for k in range(8):
        for key in ("sigma_trunk.0.linear.weight", "sigma_trunk.1.linear.bias", "geo_head.weight", "color_mlp.0.linear.weight",
                    "color_mlp.2.bias", "sigma_head.weight"):
            out[f"grad.{k}.{key}"] = g[f"submodules.{k}.{key}"].numpy()
        sub, dig = table_digest(g[f"submodules.{k}.xyz_encoder.hash_table"])
        out[f"grad.{k}.table_sub"], out[f"grad.{k}.table_digest"] = sub, dig
    for key in ("bg_mlp.0.weight", "bg_mlp.0.bias", "bg_mlp.2.weight", "bg_mlp.2.bias"):
        out[f"grad.{key}"] = g[key].numpy()
    with torch.no_grad():
        pts = (rays[:, None, :3] + rays[:, None, 3:6] * stratified_t_vals(rays[:, 6], rays[:, 7], S, False)[..., None]).reshape(-1, 3)
        w, _ = mc._routing(pts)
        out["support"] = (w > 0).numpy()
        print("  render8: rows per expert", (w > 0).sum(0).tolist(), "overlap rows", int(((w > 0).sum(1) > 1).sum()))
    save("render8.npz", **out)


def adapt_batches(steps, N):
    """Support-ray batches for the adaptation fixture: views over the whole scene, except steps 3-5, whose rays all come
    from ONE corner view (several experts then receive no ray: their parameters have grad None for those steps)."""
    wide = rays_over_scene(301, 6, 24, 24, 9.0)
    cam = dict(H=24, W=24, fx=40.0, fy=40.0, cx=12.0, cy=12.0, c2w=synth.camera_c2w(-0.7, -0.8, 0.2))
    box = SceneBox(aabb=T(synth.AABB_GLOBAL))
    dirs = get_ray_directions(24, 24, 40.0, 40.0, 12.0, 12.0, True, torch.device("cpu"))
    corner = get_rays(dirs, T(cam["c2w"]), scene_box=box).view(-1, 8)
    corner, valid = clamp_rays_near_far(corner, (None, None))
    corner = corner[valid]
    rng = np.random.default_rng(302)
    batches = []
    for it in range(steps):
        src = corner if it in (3, 4, 5) else wide
        batches.append(src[T(rng.integers(0, src.shape[0], N))])
    rays = torch.stack(batches)
    hit = rays[..., :3] + rays[..., 3:6] * ((0.45 - rays[..., 0]) / rays[..., 3])[..., None]
    gt = torch.stack([0.5 + 0.5 * torch.sin(3 * hit[..., 1]), 0.5 + 0.5 * torch.cos(2 * hit[..., 2]),
                      0.5 + 0.25 * torch.sin(2 * hit[..., 1] + hit[..., 2])], dim=-1).clamp(0, 1)
    return rays, gt


def gen_adapt():
    """BASELINE config 5: online-stage adaptation of the WHOLE container (active_module=None, 8 experts, background head)
    -- pipelines/online_stage/runtime_adapt.py:211-313 with common/utils.py get_optimizer (Adam; encoding 1e-2, sigma /
    color 2e-3, background 1e-3: configs/eval.json), clip 1.0, 16 steps, 96 samples per ray.
      fp32.*  runtime_adapt itself, called unmodified (on a CPU it switches AMP off: `enabled=... and cuda.is_available()`)
      amp.*   the same loop body with the autocast / GradScaler it enters on a GPU, on the CPU autocast policy
              (restated here line for line from :287-310, since the original hard-codes torch.cuda.amp)."""
    for name in ("lpips", "pytorch_msssim", "imageio"):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.ssim = mod.imwrite = mod.LPIPS = object
            sys.modules[name] = mod
    from common.utils import get_optimizer
    from nerfs.losses import compute_mse_loss
    from pipelines.online_stage.runtime_adapt import runtime_adapt
    out = {}
    steps, N, S = 16, 512, 96
    rays, gt = adapt_batches(steps, N)
    out["rays"], out["gt"] = rays.numpy(), gt.numpy()
    torch.manual_seed(4321)
    jit = torch.rand(steps, N, S)
    out["jitter_seed"] = np.int64(4321)
    cen = synth.CENTROIDS_G24
    boxes = synth.expert_boxes_for_grid(cen)
    conf = dict(HASH_CONF, log2_hashmap_size=12)
    P = types.SimpleNamespace(lr=1e-3, encoding_lr=1e-2, sigma_lr=2e-3, color_lr=2e-3, bg_lr=1e-3, optimizer="adam",
                              weight_decay=0.0, use_amp=True, ray_samples=S, chunk_points=1 << 22, color_space="linear")
    keep = ("submodules.0.sigma_trunk.0.linear.weight", "submodules.3.color_mlp.2.weight", "submodules.7.geo_head.bias",
            "bg_mlp.0.weight", "bg_mlp.2.bias")

    def fresh():
        m = make_container(8, cen, boxes, 1.05, True, 303, hash_conf=conf)
        with torch.no_grad():
            for sub in m.submodules:
                sub.xyz_encoder.hash_table.mul_(1e-3 / 0.5)
        m.train()
        return m

    class Loader:                      # yields (rays, rgbs) and arms the captured jitter for the step about to run
        def __init__(self):
            self.it = 0
        def __iter__(self):
            return self
        def __next__(self):
            if self.it >= steps:
                raise StopIteration
            torch.rand_like = lambda x, _j=jit[self.it]: _j
            self.it += 1
            return rays[self.it - 1], gt[self.it - 1]

    # ---- fp32: the reference function itself ----
    m = fresh()
    opt = get_optimizer(P, m)
    losses = []
    orig_rand_like = torch.rand_like
    import pipelines.online_stage.runtime_adapt as ra
    orig_loss = ra.compute_mse_loss
    def spy(*a, **k):
        l = orig_loss(*a, **k)
        losses.append(float(l.detach()))
        return l
    ra.compute_mse_loss = spy
    try:
        res = runtime_adapt(P=P, model=m, data_loader=Loader(), optimizer=opt, steps=steps, active_module=None, grad_clip=1.0)
    finally:
        torch.rand_like = orig_rand_like
        ra.compute_mse_loss = orig_loss
    assert res["steps"] == steps and len(losses) == steps
    out["fp32.loss"] = np.array(losses, np.float64)
    sd = dict(m.named_parameters())
    for k in keep:
        out[f"fp32.param.{k}"] = sd[k].detach().numpy().copy()
    out["fp32.table_sub.0"] = sd["submodules.0.xyz_encoder.hash_table"].detach().numpy()[::97].copy()
    out["fp32.adam_steps"] = np.array([float(opt.state[sd[f"submodules.{k}.sigma_head.weight"]]["step"]) if sd[f"submodules.{k}.sigma_head.weight"] in opt.state else 0.0
                                       for k in range(8)])
    print("  adapt fp32 loss first/last:", losses[0], losses[-1], "Adam steps per expert:", out["fp32.adam_steps"].tolist())

    # ---- amp: same loop with autocast(float16) + GradScaler, as on a GPU ----
    m = fresh()
    opt = get_optimizer(P, m)
    scaler = torch.amp.GradScaler("cpu")
    losses, scales = [], []
    for it in range(steps):
        opt.zero_grad()
        scales.append(scaler.get_scale())
        with torch.autocast("cpu", dtype=torch.float16):
            loss = with_jitter(jit[it], lambda: compute_mse_loss(P, model=m, data={"rays": rays[it], "rgbs": gt[it]}, params=None,
                                                                  active_module=None, reduction="mean"))
        scaler.scale(loss).backward()
        scaler.unscale_(opt)
        torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
        scaler.step(opt)
        scaler.update()
        losses.append(float(loss.detach()))
    out["amp.loss"] = np.array(losses, np.float64)
    out["amp.scales"] = np.array(scales + [scaler.get_scale()], np.float64)
    sd = dict(m.named_parameters())
    for k in keep:
        out[f"amp.param.{k}"] = sd[k].detach().numpy().copy()
    print("  adapt amp  loss first/last:", losses[0], losses[-1], "scale:", scales[0], "->", scaler.get_scale())
    save("adapt.npz", **out)


def gen_hashgrid_t19():
    """BASELINE-size table (T = 2^19, 16 levels, F = 2: 64 MiB): features of 2048 points and a digest of the table
    gradient (models/encodings.py:331-381), Linear interpolation."""
    out = {}
    P, L, F, log2T = 2048, 16, 2, 19
    x = hash_inputs(121, P)
    table = np.random.default_rng(122).uniform(-0.5, 0.5, size=(L << log2T, F)).astype(F32)
    dout = np.random.default_rng(123).standard_normal((P, L * F)).astype(F32)
    enc = HashGridEncoder(levels=L, min_res=16, max_res=4096, log2_hashmap_size=log2T, features_per_level=F, interpolation="Linear")
    with torch.no_grad():
        enc.hash_table.copy_(T(table))
    y = enc(T(x))
    (y * T(dout)).sum().backward()
    g = enc.hash_table.grad.numpy()
    out["feat"] = y.detach().numpy()
    nz = np.flatnonzero(np.abs(g).sum(1))
    keep = nz[:: max(1, len(nz) // 4096)][:4096]
    out["grad_rows"], out["grad_vals"] = keep.astype(np.int64), g[keep].copy()
    out["grad_level_sums"] = g.reshape(L, 1 << log2T, F).sum(axis=1, dtype=np.float64)
    out["grad_level_abs"] = np.abs(g).reshape(L, 1 << log2T, F).sum(axis=1, dtype=np.float64)
    out["grad_nonzero_rows"] = np.int64(len(nz))
    print("  hashgrid_t19: nonzero gradient rows", len(nz))
    save("hashgrid_t19.npz", **out)


def gen_field_half():
    """The reference's fp16 arithmetic for one expert: MetaNGP.forward under torch.autocast(float16) (CPU policy = the
    CUDA policy for these ops: matmul in fp16 with fp32 accumulation, bias add / ReLU / exp / sigmoid in fp32,
    models/metamodule/metamodule.py:142-155) on the inputs of field.npz, and its autograd gradients under a
    loss scale of 1024 (what GradScaler supplies), divided back out."""
    out = {}
    m = make_container(1, np.zeros((1, 3)), [synth.AABB_GLOBAL], 1.0, False, 31)
    ex = m.submodules[0]
    f = np.load(HERE / "field.npz")
    x6 = T(np.concatenate([f["xyz"], f["dirs"]], axis=1))
    G = T(f["G"])
    scale = 1024.0
    with torch.autocast("cpu", dtype=torch.float16):
        y = ex(x6)
    (y.float() * G * scale).sum().backward()
    out["y"] = y.detach().float().numpy()
    for key in synth.EXPERT_KEYS:
        out["grad." + key] = (dict(ex.named_parameters())[key].grad / scale).numpy()
    gt_ = ex.xyz_encoder.hash_table.grad / scale
    out["grad.table_sub"], out["grad.table_digest"] = table_digest(gt_)
    save("field_half.npz", **out)


def gen_scene_boxes():
    """The reference's SHIPPED per-expert boxes (data/drz/out/example/masks/g22_grid_bm110_ss11/scene_boxes.pt: mins, maxs,
    counts streamed over all 249 train + val images by scripts/create_clusters.py:792-973) together with what is needed to
    redo the run: every image's pose and intrinsics (the metadata files, ~100 bytes each)."""
    root = REF / "data/drz/out/example"
    mdir = root / "masks/g22_grid_bm110_ss11"
    sb = torch.load(mdir / "scene_boxes.pt")
    out = {"mins": sb["mins"].numpy(), "maxs": sb["maxs"].numpy(), "counts": sb["counts"].numpy().astype(np.int64),
           "centroids": sb["centroids"].numpy(), "aabb_global": sb["aabb_global"].numpy(),
           "margin": np.float32(sb["boundary_margin"]), "ray_samples": np.int32(sb["ray_samples"])}
    c2w, intr, hw, names = [], [], [], []
    for split in ("train", "val"):
        for mp in sorted((root / split / "metadata").glob("*.pt")):
            md = torch.load(mp, map_location="cpu")
            c2w.append(md["c2w"].numpy().astype(F32))
            intr.append(np.array([float(a) for a in md["intrinsics"]], np.float64))
            hw.append([int(md["H"]), int(md["W"])])
            names.append(f"{split}/{mp.stem}")
    out.update(c2w=np.stack(c2w), intrinsics=np.stack(intr), HW=np.array(hw, np.int32), names=np.array(names))
    print(f"  scene_boxes: {len(names)} images, counts {out['counts'].tolist()}")
    save("scene_boxes.npz", **out)


from make_golden_items import ram_rays_items  # noqa: E402


def gen_ram_rays():
    """data/ram_rays_dataset.py RamRaysDataset (single-process mode) on the items above."""
    import types as _t
    from data.ram_rays_dataset import RamRaysDataset
    items = ram_rays_items()
    mds = []
    for it in items:
        md = _t.SimpleNamespace(H=it["H"], W=it["W"], intrinsics=it["intrinsics"], c2w=T(it["c2w"]), image_index=it["image_index"], is_val=False)
        md.load_image = (lambda a=it["image"]: T(a))
        md.load_mask = (lambda m=it["mask"]: None if m is None else T(m))
        mds.append(md)
    box = SceneBox(aabb=T(synth.AABB_GLOBAL))
    out = {}
    for tag, nf in (("none", None), ("nf", (0.02, 0.5))):
        import contextlib, io as _io
        with contextlib.redirect_stdout(_io.StringIO()), contextlib.redirect_stderr(_io.StringIO()):
            ds = RamRaysDataset(mds, center_pixels=True, val_balancing=False, ray_gen_kwargs={"scene_box": box, "near_far_override": nf}, num_workers=1)
        out[f"{tag}.rgbs"], out[f"{tag}.rays"], out[f"{tag}.idx"] = ds._rgbs.numpy(), ds._rays.numpy(), ds._img_indices.numpy()
        out[f"{tag}.num_images"] = np.int32(ds._num_images)
        print(f"  ram_rays[{tag}]: {len(ds)} rays from {ds._num_images} images")
    save("ram_rays.npz", **out)


if __name__ == "__main__":
    which = set(sys.argv[1:])
    cc = None
    def want(n):
        return not which or n in which
    if want("stage1"): gen_stage1()
    if want("hashgrid"): gen_hashgrid()
    if want("field"): gen_field()
    if want("composite"): gen_composite()
    if want("routing") or want("voronoi"):
        cc = load_cc()
    if want("routing"): gen_routing(cc)
    if want("voronoi"): gen_voronoi(cc)
    if want("render"): gen_render()
    if want("train"): gen_train()
    if want("loss"): gen_loss()
    if want("optim"): gen_optim()
    if want("taskgrid"): gen_taskgrid()
    if want("train_amp"): gen_train_amp()
    if want("render8"): gen_render8()
    if want("adapt"): gen_adapt()
    if want("hashgrid_t19"): gen_hashgrid_t19()
    if want("field_half"): gen_field_half()
    if want("scene_boxes"): gen_scene_boxes()
    if want("ram_rays"): gen_ram_rays()
