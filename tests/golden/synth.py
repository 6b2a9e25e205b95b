"""Seeded synthetic inputs shared by tests/golden/make_golden.py (which runs the reference in
the build container) and by the tests (which run the oracle and the CUDA path).  Only numpy's
PCG64 is used so the numbers are identical everywhere; nothing here touches torch RNG."""
from __future__ import annotations

import numpy as np

F32 = np.float32

#: shipped global scene box (reference data/drz/out/example/masks/*/params.pt `aabb_global`)
AABB_GLOBAL = np.array([[-0.046124, -1.1, -1.1], [0.497646, 1.1, 1.1]], F32)
#: shipped 2x2 grid centroids (same file, `centroids`), cluster_2d=True
CENTROIDS_G22 = np.array(
    [[-1.1642e-10, -0.43409, -0.484345], [-1.1642e-10, -0.43409, 0.484345],
     [-1.1642e-10, 0.43409, -0.484345], [-1.1642e-10, 0.43409, 0.484345]], F32)
#: shipped per-expert boxes (scene_boxes.pt mins/maxs) for the 2x2 grid, margin 1.10
EXPERT_BOXES_G22 = np.array(
    [[[-0.046124, -1.1, -1.1], [0.497646, 0.062537, 0.06695]],
     [[-0.046124, -1.1, -0.06695], [0.497646, 0.062537, 1.1]],
     [[-0.046124, -0.062537, -1.1], [0.497646, 1.1, 0.06695]],
     [[-0.046124, -0.062537, -0.06695], [0.497646, 1.1, 1.1]]], F32)

def expert_boxes_for_grid(centroids, aabb=AABB_GLOBAL, overlap=1.10):
    """Axis-aligned per-expert boxes for a regular (y,z) grid of centroids (cluster_2d): each cell reaches half a grid
    spacing (x `overlap`) from its centroid, clipped to the global box; full x range.  Plays the role of the shipped
    scene_boxes.pt for grids that ship none (the 2x4 grid of BASELINE configs 4/5)."""
    c = np.asarray(centroids, np.float64)
    out = []
    ys, zs = np.unique(np.round(c[:, 1], 5)), np.unique(np.round(c[:, 2], 5))
    hy = 0.5 * (ys[1] - ys[0]) if len(ys) > 1 else 10.0
    hz = 0.5 * (zs[1] - zs[0]) if len(zs) > 1 else 10.0
    for k in range(c.shape[0]):
        lo = np.array([aabb[0, 0], max(aabb[0, 1], c[k, 1] - hy * overlap) if c[k, 1] - hy > aabb[0, 1] + 1e-6 and c[k, 1] > ys[0] + 1e-6 else aabb[0, 1],
                       max(aabb[0, 2], c[k, 2] - hz * overlap) if c[k, 2] > zs[0] + 1e-6 else aabb[0, 2]])
        hi = np.array([aabb[1, 0], min(aabb[1, 1], c[k, 1] + hy * overlap) if c[k, 1] < ys[-1] - 1e-6 else aabb[1, 1],
                       min(aabb[1, 2], c[k, 2] + hz * overlap) if c[k, 2] < zs[-1] - 1e-6 else aabb[1, 2]])
        out.append(np.stack([lo, hi]))
    return np.asarray(out, F32)


#: 2x4 grid of BASELINE configs 4/5: scripts/create_clusters.py:298-314 _grid_centroids(cams, 1, 2, 4, True) over the
#: synthetic camera extent y,z in [-0.9, 0.9] (the values the reference returns; also stored in routing.npz `cen8`)
CENTROIDS_G24 = np.array([[-0.04, -0.45, -0.67499995], [-0.04, -0.45, -0.22500002], [-0.04, -0.45, 0.22500002],
                          [-0.04, -0.45, 0.67499995], [-0.04, 0.44999993, -0.67499995], [-0.04, 0.44999993, -0.22500002],
                          [-0.04, 0.44999993, 0.22500002], [-0.04, 0.44999993, 0.67499995]], F32)


EXPERT_KEYS = [
    "sigma_trunk.0.linear.weight", "sigma_trunk.0.linear.bias",
    "sigma_trunk.1.linear.weight", "sigma_trunk.1.linear.bias",
    "sigma_head.weight", "sigma_head.bias", "geo_head.weight", "geo_head.bias",
    "color_mlp.0.linear.weight", "color_mlp.0.linear.bias",
    "color_mlp.1.linear.weight", "color_mlp.1.linear.bias",
    "color_mlp.2.weight", "color_mlp.2.bias",
]


def expert_shapes(E=32, H=64, G=15, C=64):
    return [(H, E), (H,), (H, H), (H,), (1, H), (1,), (G, H), (G,),
            (C, G + 16), (C,), (C, C), (C,), (3, C), (3,)]


def make_expert_params(seed, L=16, F=2, log2T=12, E=None, H=64, G=15, C=64, table_scale=0.5):
    """Random-init weights with nn.Linear-like scale, sigma bias -1 (meta_ngp.py:83-84)."""
    rng = np.random.default_rng(seed)
    E = L * F if E is None else E
    sd = {}
    for key, shp in zip(EXPERT_KEYS, expert_shapes(E, H, G, C)):
        fan_in = shp[-1] if len(shp) == 2 else None
        if key.endswith("weight"):
            bound = 1.0 / np.sqrt(fan_in)
            last_fan = fan_in
        else:
            bound = 1.0 / np.sqrt(last_fan)
        sd[key] = rng.uniform(-bound, bound, size=shp).astype(F32)
    sd["sigma_head.bias"][:] = -1.0
    sd["xyz_encoder.hash_table"] = rng.uniform(-table_scale, table_scale, size=(L << log2T, F)).astype(F32)
    return sd


def expert_weight_list(sd):
    return [sd[k] for k in EXPERT_KEYS]


def make_bg_params(seed, hidden=32):
    rng = np.random.default_rng(seed)
    b0, b1 = 1 / np.sqrt(16), 1 / np.sqrt(hidden)
    return {
        "bg_mlp.0.weight": rng.uniform(-b0, b0, (hidden, 16)).astype(F32),
        "bg_mlp.0.bias": rng.uniform(-b0, b0, (hidden,)).astype(F32),
        "bg_mlp.2.weight": rng.uniform(-b1, b1, (3, hidden)).astype(F32),
        "bg_mlp.2.bias": rng.uniform(-b1, b1, (3,)).astype(F32),
    }


def camera_c2w(y, z, yaw=0.0, x=-0.04):
    """Nadir camera (RUB -> world DRB, x = down), SURVEY 8(d)."""
    cy, sy = np.cos(yaw), np.sin(yaw)
    R0 = np.array([[0, 0, -1], [1, 0, 0], [0, -1, 0]], np.float64)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]], np.float64)  # spin about the optical axis
    R = R0 @ Rz
    c2w = np.concatenate([R, np.array([[x], [y], [z]])], axis=1)
    return c2w.astype(F32)


def random_rays_in_box(seed, N, aabb=AABB_GLOBAL):
    """Generic ray soup: origins around the box, mostly pointing inwards, a few degenerate
    (axis-parallel, zero component, missing the box)."""
    rng = np.random.default_rng(seed)
    lo, hi = aabb[0].astype(np.float64), aabb[1].astype(np.float64)
    ctr, ext = 0.5 * (lo + hi), hi - lo
    o = ctr + (rng.uniform(-1.2, 1.2, (N, 3)) * ext)
    tgt = ctr + rng.uniform(-0.5, 0.5, (N, 3)) * ext
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    k = max(1, N // 16)
    d[:k, 0] = 0.0                       # exactly zero component
    d[k:2 * k, 1] = 1e-9                 # below eps
    d[2 * k:3 * k] = np.array([1.0, 0.0, 0.0])
    d[3 * k:4 * k] *= -1.0               # pointing away
    return o.astype(F32), d.astype(F32)


def nadir_rays(seed, n_views, H=64, W=64, f=60.0):
    """Camera parameters for n_views synthetic 64x64 nadir views."""
    rng = np.random.default_rng(seed)
    cams = []
    for _ in range(n_views):
        y, z = rng.uniform(-0.9, 0.9, 2)
        yaw = rng.uniform(-np.pi, np.pi)
        cams.append(dict(H=H, W=W, fx=f, fy=f, cx=W / 2.0, cy=H / 2.0, c2w=camera_c2w(y, z, yaw)))
    return cams


def loss_inputs(seed=61, N=341):
    """Rendered linear rgb (some outside [0,1], some exactly on the clamp / sRGB thresholds) and sRGB targets."""
    rng = np.random.default_rng(seed)
    pred = rng.uniform(-0.15, 1.15, size=(N, 3)).astype(F32)
    gt = rng.uniform(-0.05, 1.05, size=(N, 3)).astype(F32)
    pred[:12, 0] = np.array([0.0, 1.0, 0.0031308, 0.0031309, 1e-4, 2e-3, 0.5, -0.0, 1.0000001, 0.99999994, 3e-3, 4e-3], F32)
    gt[:6, 1] = np.array([0.04045, 0.04046, 0.0, 1.0, 0.04, 0.05], F32)
    return pred, gt


#: parameter tensors of the optimizer fixture: (group name, shape); sizes chosen to hit the kernels' vector and tail paths
OPTIM_SHAPES = [("encoding", (515, 2)), ("sigma", (32, 16)), ("sigma", (64,)), ("sigma", (1, 64)), ("sigma", (1,)),
                ("color", (21, 31)), ("color", (3, 64)), ("color", (3,)), ("background", (32, 16)), ("background", (3,))]
#: steps whose parameters the fixture stores (2: clipped, 3: skipped by the scaler, 5: last)
OPTIM_KEEP = (2, 3, 5)
OPTIM_LRS = {"encoding": 1e-2, "sigma": 2e-3, "color": 2e-3, "background": 1e-3}


def optim_inputs(seed=67, steps=6):
    """Initial parameters and one gradient set per step; step 3 (0-based) carries a gradient that overflows under the
    loss scale, so GradScaler must skip it; step 1 has a tiny norm (no clipping), the others are clipped."""
    rng = np.random.default_rng(seed)
    params = [rng.uniform(-0.5, 0.5, size=s).astype(F32) for _, s in OPTIM_SHAPES]
    grads = []
    for it in range(steps):
        scale = 1e-3 if it == 1 else 0.3
        gs = [(rng.standard_normal(size=s) * scale).astype(F32) for _, s in OPTIM_SHAPES]
        if it == 3:
            gs[1][5, 7] = F32(3e35)
        grads.append(gs)
    return params, grads


def task_rays(seed=83, n_soup=3000, aabb=AABB_GLOBAL):
    """Packed (N,8) rays for the task-grid binning fixture: pixels of four oblique nadir views (near / far = scene-box
    slab) plus a ray soup with degenerate directions; some miss the region, some have near >= far."""
    rng = np.random.default_rng(seed)
    lo, hi = aabb[0].astype(np.float64), aabb[1].astype(np.float64)
    rays = []
    for cam in nadir_rays(seed, 4, H=40, W=40, f=32.0):
        j, i = np.meshgrid(np.arange(40), np.arange(40), indexing="ij")
        dc = np.stack([(i + 0.5 - cam["cx"]) / cam["fx"], -(j + 0.5 - cam["cy"]) / cam["fy"], -np.ones_like(i, float)], -1)
        dc /= np.linalg.norm(dc, axis=-1, keepdims=True)
        dw = dc.reshape(-1, 3) @ cam["c2w"][:, :3].astype(np.float64).T
        ow = np.broadcast_to(cam["c2w"][:, 3].astype(np.float64), dw.shape)
        rays.append(np.concatenate([ow, dw], 1))
    o, d = random_rays_in_box(seed + 1, n_soup, aabb)
    rays.append(np.concatenate([o, d], 1).astype(np.float64))
    od = np.concatenate(rays, 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        t0, t1 = (lo - od[:, :3]) / od[:, 3:], (hi - od[:, :3]) / od[:, 3:]
    with np.errstate(invalid="ignore"):
        near = np.nan_to_num(np.nanmax(np.minimum(t0, t1), 1), nan=0.0, posinf=0.0, neginf=0.0).clip(0)
        far = np.nan_to_num(np.nanmin(np.maximum(t0, t1), 1), nan=1.0, posinf=5.0, neginf=0.0)
    miss = ~(far > near) | (near > 10.0)              # rays that miss the scene box keep a short dummy segment
    near[miss], far[miss] = 0.0, 0.5
    far[::97] = near[::97] - 0.01                     # empty segments
    return np.concatenate([od, near[:, None], far[:, None]], 1).astype(F32)
