"""numpy/ctypes front-end of the CPU oracle (oracle/acn_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of acn_oracle.c.  Nothing under
adaptive_city_nerf_b200/ may import this module; only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs do.

Each wrapper mirrors one reference function (cited in the C source) on contiguous float32
numpy arrays and returns fresh arrays.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

F32 = np.float32
_fp = C.POINTER(C.c_float)


def build(force: bool = False) -> Path:
    so = _HERE / "liboracle.so"
    src = _HERE / "acn_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B", "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(str(build()))
    return _LIB


def set_threads(n: int) -> None:
    os.environ["OMP_NUM_THREADS"] = str(int(n))


def _f(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=F32)


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class FieldWeights(C.Structure):
    NAMES = ["w_t0", "b_t0", "w_t1", "b_t1", "w_sig", "b_sig", "w_geo", "b_geo",
             "w_c0", "b_c0", "w_c1", "b_c1", "w_c2", "b_c2"]
    _fields_ = [(n, C.c_void_p) for n in NAMES]


#: state_dict keys of one reference MetaNGP expert, in FieldWeights order
EXPERT_KEYS = [
    "sigma_trunk.0.linear.weight", "sigma_trunk.0.linear.bias",
    "sigma_trunk.1.linear.weight", "sigma_trunk.1.linear.bias",
    "sigma_head.weight", "sigma_head.bias", "geo_head.weight", "geo_head.bias",
    "color_mlp.0.linear.weight", "color_mlp.0.linear.bias",
    "color_mlp.1.linear.weight", "color_mlp.1.linear.bias",
    "color_mlp.2.weight", "color_mlp.2.bias",
]


def _pack(ws):
    arrs = [_f(w) for w in ws]
    st = FieldWeights(*[a.ctypes.data for a in arrs])
    return st, arrs


# ------------------------------------------------------------------ stage 1
def ray_directions(H, W, fx, fy, cx, cy, center_pixels=True):
    out = np.empty((H, W, 3), F32)
    lib().orc_ray_directions(H, W, C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy),
                             int(center_pixels), _p(out))
    return out


def aabb_intersect(o, d, aabb, eps=1e-8, max_bound=1e10, invalid=1e10):
    o, d, aabb = _f(o), _f(d), _f(aabb).reshape(6)
    N = o.shape[0]
    tmin, tmax = np.empty(N, F32), np.empty(N, F32)
    lib().orc_aabb_intersect(_p(o), _p(d), C.c_int64(N), 3, _p(aabb), C.c_float(eps),
                             C.c_float(max_bound), C.c_float(invalid), _p(tmin), _p(tmax))
    return tmin, tmax


def get_rays(dirs_cam, c2w, aabb=None, near=0.0, far=0.0, max_bound=1e10, invalid=1e10):
    d = _f(dirs_cam).reshape(-1, 3)
    c2w = _f(c2w)[:3, :4].copy()
    rays = np.empty((d.shape[0], 8), F32)
    ab = _f(aabb).reshape(6) if aabb is not None else None
    lib().orc_get_rays(_p(d), C.c_int64(d.shape[0]), _p(c2w), _p(ab) if ab is not None else None,
                       C.c_float(near), C.c_float(far), C.c_float(max_bound), C.c_float(invalid),
                       _p(rays))
    return rays


def clamp_near_far(rays, override=(None, None), eps=1e-6, invalid=float("inf")):
    rays = _f(rays).copy()
    N = rays.shape[0]
    valid = np.empty(N, np.uint8)
    has = override is not None
    n, f = (override if has else (None, None))
    lib().orc_clamp_near_far(_p(rays), C.c_int64(N), int(has),
                             C.c_float(float("nan") if n is None else n),
                             C.c_float(float("nan") if f is None else f),
                             C.c_float(eps), C.c_float(invalid), _p(valid))
    return rays, valid.astype(bool)


def linspace01(S):
    out = np.empty(S, F32)
    lib().orc_linspace01(int(S), _p(out))
    return out


def stratified_t(rays, S, jitter=None):
    rays = _f(rays)
    N = rays.shape[0]
    u = linspace01(S)
    t = np.empty((N, S), F32)
    j = _f(jitter) if jitter is not None else None
    lib().orc_stratified_t(_p(rays), C.c_int64(N), int(S), _p(u), _p(j) if j is not None else None, _p(t))
    return t


def points(rays, t):
    rays, t = _f(rays), _f(t)
    N, S = t.shape
    pts = np.empty((N, S, 3), F32)
    lib().orc_points(_p(rays), C.c_int64(N), int(S), _p(t), _p(pts))
    return pts


# ------------------------------------------------------------------ stage 2
def world_to_unit(x, box_min, extent):
    x = _f(x).reshape(-1, 3)
    out = np.empty_like(x)
    lib().orc_world_to_unit(_p(x), C.c_int64(x.shape[0]), _p(_f(box_min)), _p(_f(extent)), _p(out))
    return out


def level_resolutions(L=16, min_res=16, max_res=4096):
    res = np.empty(L, np.int32)
    lib().orc_level_resolutions(int(L), int(min_res), int(max_res), _p(res))
    return res


INTERP = {"Nearest": 0, "Linear": 1, "Smoothstep": 2, None: 1}


def hashgrid_fwd(x01, table, L, F, log2T, res, interp="Linear", want_idx=False):
    x01, table = _f(x01).reshape(-1, 3), _f(table)
    P = x01.shape[0]
    res = np.ascontiguousarray(res, np.int32)
    out = np.empty((P, L * F), F32)
    idx = np.zeros((P, L, 8), np.int32) if want_idx else None
    lib().orc_hashgrid_fwd(_p(x01), C.c_int64(P), _p(table), int(L), int(F), int(log2T), _p(res),
                           INTERP[interp], _p(out), _p(idx) if want_idx else None)
    return (out, idx) if want_idx else out


def hashgrid_bwd(x01, dout, L, F, log2T, res, interp="Linear"):
    x01, dout = _f(x01).reshape(-1, 3), _f(dout)
    P = x01.shape[0]
    res = np.ascontiguousarray(res, np.int32)
    dtable = np.zeros((L << log2T, F), F32)
    lib().orc_hashgrid_bwd(_p(x01), C.c_int64(P), _p(dout), int(L), int(F), int(log2T), _p(res),
                           INTERP[interp], _p(dtable))
    return dtable


# ------------------------------------------------------------------ stage 3
def sh16(d):
    d = _f(d).reshape(-1, 3)
    out = np.empty((d.shape[0], 16), F32)
    lib().orc_sh16(_p(d), C.c_int64(d.shape[0]), _p(out))
    return out


def _dims(ws):
    H, E = ws[0].shape
    G = ws[6].shape[0]
    Cc = ws[8].shape[0]
    assert ws[8].shape[1] == G + 16
    return E, H, G, Cc


def field_fwd(enc, dirs, ws, half=False):
    """ws: 14 arrays in EXPERT_KEYS order.  Returns rgb_sigma (P,4)."""
    enc, dirs = _f(enc), _f(dirs).reshape(-1, 3)
    E, H, G, Cc = _dims(ws)
    st, keep = _pack(ws)
    P = enc.shape[0]
    out = np.empty((P, 4), F32)
    lib().orc_field_fwd(_p(enc), _p(dirs), C.c_int64(P), E, H, G, Cc, C.byref(st), int(half), _p(out), None)
    return out


def field_bwd(enc, dirs, ws, d_rgb_sigma):
    """Returns (list of 14 grads, d_enc)."""
    enc, dirs, dy = _f(enc), _f(dirs).reshape(-1, 3), _f(d_rgb_sigma)
    E, H, G, Cc = _dims(ws)
    st, keep = _pack(ws)
    grads = [np.zeros_like(_f(w)) for w in ws]
    gs = FieldWeights(*[g.ctypes.data for g in grads])
    P = enc.shape[0]
    d_enc = np.empty((P, E), F32)
    lib().orc_field_bwd(_p(enc), _p(dirs), C.c_int64(P), E, H, G, Cc, C.byref(st), _p(dy), C.byref(gs), _p(d_enc))
    return grads, d_enc


def background(dirs, w0, b0, w1, b1):
    dirs = _f(dirs).reshape(-1, 3)
    w0, b0, w1, b1 = _f(w0), _f(b0), _f(w1), _f(b1)
    out = np.empty((dirs.shape[0], 3), F32)
    lib().orc_background(_p(dirs), C.c_int64(dirs.shape[0]), int(w0.shape[0]), _p(w0), _p(b0), _p(w1), _p(b1), _p(out))
    return out


# ------------------------------------------------------------------ stage 4
def composite_fwd(rgb_sigma, t, bg=None, sigma_scale=1.0):
    rs, t = _f(rgb_sigma), _f(t)
    N, S = t.shape
    bgp = _f(bg) if bg is not None else None
    rgb, dep, w, acc = np.empty((N, 3), F32), np.empty(N, F32), np.empty((N, S), F32), np.empty(N, F32)
    lib().orc_composite_fwd(_p(rs), _p(t), _p(bgp) if bgp is not None else None, C.c_int64(N), int(S),
                            C.c_float(sigma_scale), _p(rgb), _p(dep), _p(w), _p(acc))
    return rgb, dep, w, acc


def composite_bwd(rgb_sigma, t, bg, g_rgb=None, g_depth=None, g_weights=None, g_acc=None, sigma_scale=1.0):
    rs, t = _f(rgb_sigma), _f(t)
    N, S = t.shape
    bgp = _f(bg) if bg is not None else None
    opt = lambda a: (_f(a) if a is not None else None)
    g_rgb, g_depth, g_weights, g_acc = opt(g_rgb), opt(g_depth), opt(g_weights), opt(g_acc)
    pp = lambda a: (_p(a) if a is not None else None)
    d = np.empty((N, S, 4), F32)
    d_bg = np.zeros((N, 3), F32) if bgp is not None else None
    lib().orc_composite_bwd(_p(rs), _p(t), pp(bgp), C.c_int64(N), int(S), C.c_float(sigma_scale),
                            pp(g_rgb), pp(g_depth), pp(g_weights), pp(g_acc), _p(d), pp(d_bg))
    return d, d_bg


# ------------------------------------------------------------------ stage 5
def route_points(pts, centroids, margin, cluster_2d=True):
    pts, cen = _f(pts), _f(centroids)
    P, K = pts.shape[0], cen.shape[0]
    dims = 2 if cluster_2d else 3
    if margin > 1.0:
        w = np.empty((P, K), F32)
        lib().orc_route_points(_p(pts), C.c_int64(P), int(pts.shape[1]), _p(cen), K, dims, C.c_float(margin), _p(w), None)
        return w, None
    hard = np.empty(P, np.int32)
    lib().orc_route_points(_p(pts), C.c_int64(P), int(pts.shape[1]), _p(cen), K, dims, C.c_float(margin), None, _p(hard))
    return None, hard


def route_rays_voronoi(rays, S, centroids, margin, cluster_2d=True):
    rays, cen = _f(rays), _f(centroids)
    N, K = rays.shape[0], cen.shape[0]
    u = linspace01(S)
    mask = np.empty((N, K), np.uint8)
    lib().orc_route_rays_voronoi(_p(rays), C.c_int64(N), int(S), _p(u), _p(cen), K, 2 if cluster_2d else 3,
                                 C.c_float(margin), _p(mask))
    return mask.astype(bool)


def blend(y_all, weights=None, hard=None):
    y = _f(y_all)
    K, P, _ = y.shape
    out = np.empty((P, 4), F32)
    w = _f(weights) if weights is not None else None
    h = np.ascontiguousarray(hard, np.int32) if hard is not None else None
    lib().orc_blend(_p(y), C.c_int64(P), K, _p(w) if w is not None else None, _p(h) if h is not None else None, _p(out))
    return out


# ------------------------------------------------------------------ glue (one expert)
def render_expert(rays, S, ws, table, box_min, extent, L, F, log2T, res, jitter=None, bg=None,
                  half=False, interp="Linear"):
    """nerfs/ray_rendering.py:290-345 render_rays_stratified with active_module set."""
    rays = _f(rays)
    t = stratified_t(rays, S, jitter)
    pts = points(rays, t).reshape(-1, 3)
    x01 = world_to_unit(pts, box_min, extent)
    enc = hashgrid_fwd(x01, table, L, F, log2T, res, interp)
    dirs = np.repeat(rays[:, 3:6], S, axis=0)
    rs = field_fwd(enc, dirs, ws, half=half)
    out = composite_fwd(rs.reshape(rays.shape[0], S, 4), t, bg)
    return out + (dict(t=t, x01=x01, enc=enc, dirs=dirs, rgb_sigma=rs),)


COLOR_SPACE = {"linear": 0, "srgb": 1, "identity": 2}


def color_mse(pred, gt, color_space="linear", reduction="mean"):
    """color_space_transformer + F.mse_loss -> (loss or per-element squared errors, d loss / d pred)."""
    p, g = _f(pred), _f(np.broadcast_to(gt, np.shape(pred)))
    elem, dpred = np.empty_like(p), np.empty_like(p)
    fn = lib().orc_color_mse
    fn.restype = C.c_double
    s = fn(_p(p), _p(g), C.c_int64(p.size), COLOR_SPACE[color_space], _p(elem), _p(dpred))
    if reduction == "none":
        return elem, dpred
    if reduction == "sum":
        return F32(s), dpred
    return F32(s / p.size) if p.size else F32(np.nan), (dpred / F32(p.size)).astype(F32)


class AdamState:
    """Parameters, moments and step count of orc_adam_step (arrays are updated in place)."""

    def __init__(self, params, lrs, wds=None, betas=(0.9, 0.999), eps=1e-8, adamw=False):
        self.p = [_f(a).copy() for a in params]
        self.m = [np.zeros_like(a) for a in self.p]
        self.v = [np.zeros_like(a) for a in self.p]
        self.lr = np.asarray(lrs, np.float64)
        self.wd = np.zeros(len(self.p)) if wds is None else np.asarray(wds, np.float64)
        self.betas, self.eps, self.adamw = betas, eps, adamw
        self.step = C.c_double(0.0)

    def update(self, grads, grad_scale=1.0, max_norm=0.0):
        """-> (total_norm, skipped); returns the unscaled, clipped gradients in `self.g`."""
        T = len(self.p)
        self.g = [_f(a).copy() for a in grads]
        arr = lambda xs: (C.c_void_p * T)(*[x.ctypes.data for x in xs])
        n = (C.c_int64 * T)(*[x.size for x in self.p])
        skipped = C.c_int(0)
        fn = lib().orc_adam_step
        fn.restype = C.c_float
        norm = fn(T, arr(self.p), arr(self.g), arr(self.m), arr(self.v), n, _p(self.lr), _p(self.wd),
                  C.c_double(self.betas[0]), C.c_double(self.betas[1]), C.c_double(self.eps), int(self.adamw),
                  C.c_float(grad_scale), C.c_float(max_norm or 0.0), C.byref(self.step), C.byref(skipped))
        return float(norm), bool(skipped.value)


def dda_route_rays(rays, aabb, cells, cell3, cell_bounds, tol, max_steps=64):
    """data/task_dataset.py _route_and_bin, "dda" policy -> (cell id per ray or -1, best in-cell length, counts (C,))."""
    rays = _f(rays)
    N = rays.shape[0]
    C_ = int(np.prod(cells))
    cid = np.empty(N, np.int32)
    blen = np.empty(N, F32)
    counts = np.zeros(C_, np.int64)
    cells_a = np.asarray(cells, np.int32)
    lib().orc_dda_route_rays(_p(rays), C.c_int64(N), _p(_f(aabb).reshape(-1)), _p(cells_a), _p(_f(cell3)),
                             _p(_f(cell_bounds).reshape(-1)), _p(_f(tol)), int(max_steps), _p(cid), _p(blen), _p(counts))
    return cid, blen, counts
