"""Operator layer: one thin Python function per C-ABI entry point plus the torch.autograd
Functions that stitch the kernels into PyTorch's graph.  Python owns all storage; the shims
make inputs contiguous fp32, allocate outputs with torch.empty and pass raw pointers + the
current CUDA stream.  Nothing here computes on the CPU."""
from __future__ import annotations

import ctypes as C_
import os
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor
from torch.autograd.function import once_differentiable

from . import _lib
from ._lib import F16, F32, check, ctx, dev_f32, lib, pack_weights, ptr, stream

_NAN = float("nan")


def _dt(t: Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.float16:
        return F16
    raise RuntimeError(f"unsupported dtype {t.dtype}")


# ------------------------------------------------------------------------------------------ stage 1
def ray_directions(H: int, W: int, fx: float, fy: float, cx: float, cy: float, center_pixels: bool,
                   device: torch.device) -> Tensor:
    device = torch.device(device)
    out = torch.empty(H, W, 3, dtype=torch.float32, device=device)
    check(lib().acn_ray_directions(ctx(device), H, W, fx, fy, cx, cy, int(bool(center_pixels)), ptr(out), stream(device)))
    return out


def aabb_intersect(o: Tensor, d: Tensor, aabb: Tensor, eps=1e-8, max_bound=1e10, invalid=1e10) -> Tuple[Tensor, Tensor]:
    o, d = dev_f32(o, "origins"), dev_f32(d, "directions")
    aabb = dev_f32(aabb.to(o.device), "aabb").reshape(6)
    N = o.shape[0]
    tmin = torch.empty(N, dtype=torch.float32, device=o.device)
    tmax = torch.empty_like(tmin)
    check(lib().acn_aabb_intersect(ctx(o.device), ptr(o), ptr(d), N, o.stride(0) if N else 3, d.stride(0) if N else 3,
                                   ptr(aabb), eps, max_bound, invalid, ptr(tmin), ptr(tmax), stream(o.device)))
    return tmin, tmax


def get_rays(dirs_cam: Tensor, c2w: Tensor, aabb: Optional[Tensor], near: float, far: float,
             max_bound: float, invalid: float) -> Tensor:
    d = dev_f32(dirs_cam, "directions").reshape(-1, 3)
    c = dev_f32(c2w.to(d.device), "c2w")
    assert c.shape[-1] == 4 and c.shape[0] in (3, 4), "c2w must be (3,4) or (4,4)"
    a = dev_f32(aabb.to(d.device), "aabb").reshape(6) if aabb is not None else None
    N = d.shape[0]
    rays = torch.empty(N, 8, dtype=torch.float32, device=d.device)
    check(lib().acn_get_rays(ctx(d.device), ptr(d), N, ptr(c), ptr(a), float(near), float(far), max_bound, invalid,
                             ptr(rays), stream(d.device)))
    return rays


def clamp_near_far_(rays: Tensor, has_override: bool, n: Optional[float], f: Optional[float], eps: float,
                    invalid: float) -> Tensor:
    """In place on `rays` (N,8) contiguous fp32; returns the bool validity mask."""
    assert rays.is_contiguous() and rays.dtype == torch.float32
    N = rays.shape[0]
    valid = torch.empty(N, dtype=torch.uint8, device=rays.device)
    check(lib().acn_clamp_near_far(ctx(rays.device), ptr(rays), N, int(has_override),
                                   _NAN if n is None else float(n), _NAN if f is None else float(f),
                                   eps, invalid, ptr(valid), stream(rays.device)))
    return valid.bool()


_ULIN = {}


def u_lin(S: int, device: torch.device) -> Tensor:
    """torch.linspace(0,1,S) evaluated by the CPU kernel the reference oracle uses (bit-exact
    bins), cached per (S, device)."""
    key = (int(S), str(device))
    t = _ULIN.get(key)
    if t is None:
        t = torch.linspace(0.0, 1.0, int(S), device="cpu").to(device)
        _ULIN[key] = t
    return t


def sample_stratified(rays: Tensor, S: int, jitter: Optional[Tensor]) -> Tensor:
    rays = dev_f32(rays, "rays")
    N = rays.shape[0]
    t = torch.empty(N, S, dtype=torch.float32, device=rays.device)
    j = dev_f32(jitter, "jitter") if jitter is not None else None
    check(lib().acn_sample_stratified(ctx(rays.device), ptr(rays), N, S, ptr(u_lin(S, rays.device)), ptr(j), ptr(t),
                                      stream(rays.device)))
    return t


def points(rays: Tensor, t: Tensor) -> Tensor:
    rays, t = dev_f32(rays, "rays"), dev_f32(t, "t_vals")
    N, S = t.shape
    id6 = torch.empty(N * S, 6, dtype=torch.float32, device=rays.device)
    check(lib().acn_points(ctx(rays.device), ptr(rays), N, S, ptr(t), ptr(id6), stream(rays.device)))
    return id6


def rays_coherent_flag(rays: Tensor, S: int, threshold: float = 0.25) -> Tensor:
    """Device int32 flag: 1 when consecutive rays are neighbouring pixels of a frame (the gap between ray r and r+1 at mid
    depth is well below the spacing of the samples along a ray).  The gather kernels then give a warp one sample of 32
    adjacent rays (`ray_major`) instead of 32 samples of one ray; handing them this tensor keeps the choice on the
    device (no host read).  A performance hint only: results do not depend on it."""
    rays = dev_f32(rays, "rays")
    flag = torch.empty(1, dtype=torch.int32, device=rays.device)
    check(lib().acn_rays_coherent(ctx(rays.device), ptr(rays), rays.shape[0], int(S), float(threshold), ptr(flag),
                                  stream(rays.device)))
    return flag


def rays_are_coherent(rays: Tensor, S: int, threshold: float = 0.25) -> bool:
    """`rays_coherent_flag` read back to the host."""
    return bool(rays_coherent_flag(rays, S, threshold).item())


def _ray_major_args(ray_major):
    """(int, device pointer) for the `ray_major` / `ray_major_dev_or_null` pair of the C ABI."""
    if isinstance(ray_major, Tensor):
        return 0, ptr(ray_major)
    return int(bool(ray_major)), None


# ------------------------------------------------------------------------------------------ stage 2
class GridSpec:
    """Static description of one hash grid (device-resident resolution table)."""

    def __init__(self, L: int, F: int, log2T: int, res: Tensor, interp: int, res_host: Optional[Sequence[int]] = None):
        self.L, self.F, self.log2T, self.res, self.interp = L, F, log2T, res, interp
        #: the same resolutions as plain host ints (known at construction, so no device read): lets the fused forward
        #: size the shared-memory lattices of the coarsest levels
        self.res_host = (C_.c_int32 * len(res_host))(*[int(r) for r in res_host]) if res_host is not None else None

    @property
    def rows(self) -> int:
        return self.L << self.log2T


def _grid_res(spec: GridSpec, device) -> Tensor:
    if spec.res.device != device or spec.res.dtype != torch.int32:
        spec.res = spec.res.to(device=device, dtype=torch.int32).contiguous()
    return spec.res


def hashgrid_fwd(x: Tensor, table: Tensor, spec: GridSpec, box6: Optional[Tensor] = None,
                 out_dtype=torch.float32, want_idx: bool = False):
    """x (P,>=3) fp32 (row stride honoured) -> (P, L*F); optional (P,L,8) int32 table rows."""
    assert x.dim() == 2 and x.shape[1] >= 3
    if not x.is_cuda:
        raise RuntimeError("hashgrid_fwd needs CUDA tensors; there is no CPU path")
    if x.dtype != torch.float32 or x.stride(1) != 1:
        x = x.float().contiguous()
    P = x.shape[0]
    dev = x.device
    table = dev_f32(table, "hash_table")
    out = torch.empty(P, spec.L * spec.F, dtype=out_dtype, device=dev)
    idx = torch.zeros(P, spec.L, 8, dtype=torch.int32, device=dev) if want_idx else None
    check(lib().acn_hashgrid_fwd(ctx(dev), ptr(x), P, x.stride(0) if P else 3, None, ptr(box6), ptr(table), spec.L, spec.F,
                                 spec.log2T, ptr(_grid_res(spec, dev)), spec.interp, ptr(out), _dt(out), ptr(idx),
                                 stream(dev)))
    return (out, idx) if want_idx else out


def hashgrid_bwd(x: Tensor, dout: Tensor, spec: GridSpec, box6: Optional[Tensor], dtable: Tensor) -> None:
    if x.dtype != torch.float32 or x.stride(1) != 1:
        x = x.float().contiguous()
    dout = dout.contiguous()
    P = x.shape[0]
    dev = x.device
    check(lib().acn_hashgrid_bwd(ctx(dev), ptr(x), P, x.stride(0) if P else 3, None, ptr(box6), spec.L, spec.F, spec.log2T,
                                 ptr(_grid_res(spec, dev)), spec.interp, ptr(dout), _dt(dout), ptr(dtable), stream(dev)))


def hashgrid_fwd_rays(rays: Tensor, t: Tensor, table: Tensor, spec: GridSpec, box6: Optional[Tensor], out_dtype,
                      ray_major=False) -> Tensor:
    N, S = t.shape
    dev = rays.device
    out = torch.empty(N * S, spec.L * spec.F, dtype=out_dtype, device=dev)
    check(lib().acn_hashgrid_fwd_rays(ctx(dev), ptr(rays), ptr(t), N, S, ptr(box6), ptr(table), spec.L, spec.F, spec.log2T,
                                      ptr(_grid_res(spec, dev)), spec.interp, ptr(out), _dt(out), *_ray_major_args(ray_major), stream(dev)))
    return out


def hashgrid_bwd_rays(rays: Tensor, t: Tensor, dout: Tensor, spec: GridSpec, box6: Optional[Tensor], dtable: Tensor) -> None:
    N, S = t.shape
    dev = rays.device
    dout = dout.contiguous()
    check(lib().acn_hashgrid_bwd_rays(ctx(dev), ptr(rays), ptr(t), N, S, ptr(box6), spec.L, spec.F, spec.log2T,
                                      ptr(_grid_res(spec, dev)), spec.interp, ptr(dout), _dt(dout), ptr(dtable), stream(dev)))


class HashEncodeFn(torch.autograd.Function):
    """models/encodings.py:293-381 forward + table gradient (positions carry no gradient, as in
    the reference pipeline where rays and t are built under no_grad)."""

    @staticmethod
    def forward(ctx_, x, table, spec, box6):
        x2 = x.reshape(-1, x.shape[-1])
        out = hashgrid_fwd(x2, table, spec, box6)
        ctx_.save_for_backward(x2, box6 if box6 is not None else x2.new_empty(0))
        ctx_.spec, ctx_.has_box, ctx_.tshape = spec, box6 is not None, table.shape
        return out.view(*x.shape[:-1], spec.L * spec.F)

    @staticmethod
    @once_differentiable
    def backward(ctx_, g):
        x2, box6 = ctx_.saved_tensors
        if ctx_.needs_input_grad[0]:
            raise NotImplementedError("gradient w.r.t. hash-grid input positions is not implemented (the reference never needs it)")
        dtable = None
        if ctx_.needs_input_grad[1]:
            dtable = torch.zeros(ctx_.tshape, dtype=torch.float32, device=g.device)
            hashgrid_bwd(x2, g.reshape(x2.shape[0], -1).float(), ctx_.spec, box6 if ctx_.has_box else None, dtable)
        return None, dtable, None, None


# ------------------------------------------------------------------------------------------ stage 3
def sh16(d: Tensor) -> Tensor:
    d2 = dev_f32(d, "directions").reshape(-1, 3)
    out = torch.empty(d2.shape[0], 16, dtype=torch.float32, device=d2.device)
    check(lib().acn_sh16(ctx(d2.device), ptr(d2), d2.shape[0], 3, ptr(out), stream(d2.device)))
    return out.view(*d.shape[:-1], 16)


def _field_dims(ws: Sequence[Tensor]):
    H, E = ws[0].shape
    G = ws[6].shape[0]
    C = ws[8].shape[0]
    return int(E), int(H), int(G), int(C)


def field_fwd(enc: Tensor, dirs: Tensor, dirs_stride: int, dirs_group: int, ws: Sequence[Tensor], half: bool) -> Tensor:
    P = enc.shape[0]
    E, H, G, C = _field_dims(ws)
    dev = enc.device
    if half and enc.dtype != torch.float16:
        enc = enc.half()                       # the tensor-core kernels take fp16 encodings
    enc = enc.contiguous()
    out = torch.empty(P, 4, dtype=torch.float32, device=dev)
    st = pack_weights(ws)
    check(lib().acn_field_fwd(ctx(dev), ptr(enc), _dt(enc), ptr(dirs), dirs_stride, dirs_group, P, None, E, H, G, C, C_.byref(st),
                              F16 if half else F32, ptr(out), stream(dev)))
    return out


def field_bwd(enc: Tensor, dirs: Tensor, dirs_stride: int, dirs_group: int, ws: Sequence[Tensor], half: bool,
              d_rgb_sigma: Tensor, want_enc_grad: bool, need: Sequence[bool]):
    P = enc.shape[0]
    E, H, G, C = _field_dims(ws)
    dev = enc.device
    if half and enc.dtype != torch.float16:
        enc = enc.half()
    enc = enc.contiguous()
    # one zero-filled buffer behind all the gradient tensors: one memset instead of up to 14 fill launches
    flat = torch.zeros(sum(w.numel() for w, n in zip(ws, need) if n), dtype=torch.float32, device=dev)
    grads: List[Optional[Tensor]] = []
    off = 0
    for w, n in zip(ws, need):
        grads.append(flat[off:off + w.numel()].view(w.shape) if n else None)
        off += w.numel() if n else 0
    # d_enc stays fp32 even for an fp16 encoding: per-sample gradients of a mean loss over 2^18 rays are
    # ~1e-9 and would flush to zero in fp16 (the reference needs GradScaler for the same reason)
    d_enc = torch.empty(P, E, dtype=torch.float32, device=dev) if want_enc_grad else None
    wst, gst = pack_weights(ws), pack_weights(grads)
    check(lib().acn_field_bwd(ctx(dev), ptr(enc), _dt(enc), ptr(dirs), dirs_stride, dirs_group, P, None, E, H, G, C,
                              C_.byref(wst), F16 if half else F32, ptr(d_rgb_sigma), C_.byref(gst),
                              ptr(d_enc), F32, stream(dev)))
    return grads, d_enc


#: set False to run the two-kernel forward (hash encode -> fp16 encoding in HBM -> MLP); tests cross-check the two
FUSED_EXPERT_FWD = os.environ.get("ACN_FUSED_FWD", "1") != "0"
#: stage the coarsest hash levels as dense lattices in shared memory inside the fused forward.  OFF: measured 7.2 ms vs
#: 3.76 ms per 2^24 samples -- the 150 KB of lattices come out of the unified L1/shared-memory array, and the L1 that is
#: left no longer holds the mid levels' rows
STAGE_COARSE_LEVELS = os.environ.get("ACN_STAGE_COARSE", "0") != "0"


def fused_fwd_ok(spec: "GridSpec", ws: Sequence[Tensor], half: bool) -> bool:
    E, H, G, C = _field_dims(ws)
    return (FUSED_EXPERT_FWD and half and spec.F == 2 and spec.L in (8, 16) and spec.interp != 0 and H == 64 and C == 64
            and 1 <= G <= 15 and E == spec.L * spec.F)


def render_expert_fwd(pos: Sequence[Tensor], table: Tensor, spec: "GridSpec", box6: Optional[Tensor], dirs: Tensor,
                      dirs_stride: int, dirs_group: int, ws: Sequence[Tensor], want_enc: bool, ray_major=False,
                      rng: Optional[Tensor] = None, enc: Optional[Tensor] = None, out: Optional[Tensor] = None):
    """Fused hash encode + tcgen05 MLPs (acn_render_expert_fwd) -> (rgb_sigma (P,4) fp32, enc (P,E) fp16 or None).
    pos: (rays (N,8), t (N,S)) or (x (P,>=3),); rng: device int32 [first, end) row range (buckets)."""
    _, H, G, C = _field_dims(ws)
    if len(pos) == 2:
        rays, t = pos
        x, xs, S, P = None, 3, t.shape[1], t.shape[0] * t.shape[1]
    else:
        x, xs, rays, t, S, P = pos[0], pos[0].stride(0), None, None, 1, pos[0].shape[0]
    dev = dirs.device
    if want_enc and enc is None:
        enc = torch.empty(P, spec.L * spec.F, dtype=torch.float16, device=dev)
    if out is None:
        out = torch.empty(P, 4, dtype=torch.float32, device=dev)
    rm, rm_dev = _ray_major_args(ray_major)
    st = pack_weights(ws)
    res_host = spec.res_host if (STAGE_COARSE_LEVELS and P >= 65536) else None
    check(lib().acn_render_expert_fwd(ctx(dev), ptr(x), xs, ptr(rays), ptr(t), P, S, rm, rm_dev, ptr(rng), ptr(box6), ptr(table),
                                      spec.L, spec.F, spec.log2T, ptr(_grid_res(spec, dev)), res_host, spec.interp, ptr(dirs),
                                      dirs_stride, dirs_group, H, G, C, C_.byref(st), ptr(enc) if want_enc else None, ptr(out),
                                      stream(dev)))
    return out, (enc if want_enc else None)


def render_expert_bwd(enc: Tensor, pos: Sequence[Tensor], dirs: Tensor, dirs_stride: int, dirs_group: int, ws: Sequence[Tensor],
                      d_rgb_sigma: Tensor, need: Sequence[bool], spec: "GridSpec", box6: Optional[Tensor], dtable: Tensor):
    """Fused MLP backward + table scatter (acn_render_expert_bwd): -> the 14 weight gradients; dtable is accumulated into."""
    P = enc.shape[0]
    E, H, G, C = _field_dims(ws)
    dev = enc.device
    flat = torch.zeros(sum(w.numel() for w, n in zip(ws, need) if n), dtype=torch.float32, device=dev)
    grads: List[Optional[Tensor]] = []
    off = 0
    for w, n in zip(ws, need):
        grads.append(flat[off:off + w.numel()].view(w.shape) if n else None)
        off += w.numel() if n else 0
    wst, gst = pack_weights(ws), pack_weights(grads)
    if len(pos) == 2:
        rays, t = pos
        x, xs, S = None, 3, t.shape[1]
    else:
        x, xs, rays, t, S = pos[0], pos[0].stride(0), None, None, 1
    check(lib().acn_render_expert_bwd(ctx(dev), ptr(x), xs, ptr(rays), ptr(t), P, S, None, ptr(box6), spec.L, spec.F, spec.log2T,
                                      ptr(_grid_res(spec, dev)), spec.interp, ptr(enc), ptr(dirs), dirs_stride, dirs_group,
                                      H, G, C, C_.byref(wst), ptr(d_rgb_sigma), C_.byref(gst), ptr(dtable), stream(dev)))
    return grads


def debug_render_expert_bwd_single(enc: Tensor, pos: Sequence[Tensor], dirs: Tensor, dirs_stride: int, dirs_group: int,
                                   ws: Sequence[Tensor], d_rgb_sigma: Tensor, need: Sequence[bool], spec: "GridSpec",
                                   box6: Optional[Tensor], dtable: Tensor):
    """The single-role fused backward of the debug library (acn_debug_render_expert_bwd_single): cross-check / A-B partner
    of render_expert_bwd; same arguments and results."""
    l, chk, dctx = _dbg()
    P = enc.shape[0]
    E, H, G, C = _field_dims(ws)
    dev = enc.device
    flat = torch.zeros(sum(w.numel() for w, n in zip(ws, need) if n), dtype=torch.float32, device=dev)
    grads: List[Optional[Tensor]] = []
    off = 0
    for w, n in zip(ws, need):
        grads.append(flat[off:off + w.numel()].view(w.shape) if n else None)
        off += w.numel() if n else 0
    wst, gst = pack_weights(ws), pack_weights(grads)
    if len(pos) == 2:
        rays, t = pos
        x, xs, S = None, 3, t.shape[1]
    else:
        x, xs, rays, t, S = pos[0], pos[0].stride(0), None, None, 1
    chk(l.acn_debug_render_expert_bwd_single(dctx(dev), ptr(x), xs, ptr(rays), ptr(t), P, S, None, ptr(box6), spec.L, spec.F, spec.log2T,
                                             ptr(_grid_res(spec, dev)), spec.interp, ptr(enc), ptr(dirs), dirs_stride, dirs_group,
                                             H, G, C, C_.byref(wst), ptr(d_rgb_sigma), C_.byref(gst), ptr(dtable), stream(dev)))
    return grads


#: set False to run the two-kernel backward (MLP backward -> d_enc in HBM -> scatter); tests cross-check the two
FUSED_EXPERT_BWD = True


class ExpertFieldFn(torch.autograd.Function):
    """One expert on a batch of points: world->unit, hash encode, density trunk + heads, SH,
    colour MLP, activations (models/inr/meta_ngp.py:226-241) -> (P,4) [rgb, sigma].

    Positions come either from `x6` (P,>=6: xyz + dir) or from (`rays` (N,8), `t` (N,S)).
    Differentiable inputs: the hash table and the 14 MLP tensors (module parameters or `params=`
    fast weights).  The backward recomputes the MLP activations from the saved encoding."""

    @staticmethod
    def forward(ctx_, x6, rays, t, table, spec, box6, half, ray_major, table_node, *ws):
        ws = [dev_f32(w, "MLP weight") for w in ws]
        table_c = dev_f32(table, "hash_table")
        enc_dtype = torch.float16 if half else torch.float32
        fused = fused_fwd_ok(spec, ws, half)
        # grad mode is off inside forward(): the caller's mode rides in with the node (grad_ctx)
        table_node, grad_on = table_node if isinstance(table_node, tuple) else (table_node, True)
        want_enc = grad_on and any(ctx_.needs_input_grad)          # inference never materialises the encoding
        if rays is not None:
            rays = dev_f32(rays, "rays")
            t = dev_f32(t, "t_vals")
            dirs, dstride, dgroup = rays[:, 3:], 8, t.shape[1]
            pos = (rays, t)
            if not fused:
                enc = hashgrid_fwd_rays(rays, t, table_c, spec, box6, enc_dtype, ray_major)
        else:
            if x6.dtype != torch.float32 or not x6.is_contiguous():
                x6 = x6.float().contiguous()
            dirs, dstride, dgroup = x6[:, 3:], x6.stride(0), 1
            pos = (x6,)
            if not fused:
                enc = hashgrid_fwd(x6, table_c, spec, box6, enc_dtype)
        if fused:
            out, enc = render_expert_fwd(pos, table_c, spec, box6, dirs, dstride, dgroup, ws, want_enc, ray_major)
            if enc is None:
                enc = out.new_empty(0)
        else:
            out = field_fwd(enc, dirs, dstride, dgroup, ws, half)
        ctx_.save_for_backward(enc, *pos, *(t_ for t_ in (box6,) if t_ is not None), *ws)
        ctx_.meta = (spec, half, len(pos), box6 is not None, dstride, dgroup, table.shape)
        ctx_.table_node = table_node
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx_, g):
        spec, half, npos, has_box, dstride, dgroup, tshape = ctx_.meta
        saved = ctx_.saved_tensors
        enc, pos = saved[0], saved[1:1 + npos]
        box6 = saved[1 + npos] if has_box else None
        ws = saved[1 + npos + int(has_box):]
        # needs_input_grad is static (the table is a Parameter); what THIS backward pass wants is not: the inner loop
        # asks torch.autograd.grad for the 14 fast weights only (meta_core.py:54-59) and must not pay for d_enc and the
        # table scatter -- the reference's un-fused graph gets that pruning from the engine for free.
        need_table = ctx_.needs_input_grad[3] and engine_wants(ctx_.table_node)
        need_w = ctx_.needs_input_grad[9:]
        dirs = pos[0][:, 3:]
        g = g.contiguous().float()
        if (FUSED_EXPERT_BWD and need_table and half and spec.F == 2 and spec.L in (8, 16) and spec.interp != 0
                and enc.dtype == torch.float16):
            dtable = torch.zeros(tshape, dtype=torch.float32, device=g.device)
            grads = render_expert_bwd(enc, pos, dirs, dstride, dgroup, ws, g, need_w, spec, box6, dtable)
            return (None, None, None, dtable, None, None, None, None, None, *grads)
        grads, d_enc = field_bwd(enc, dirs, dstride, dgroup, ws, half, g, need_table, need_w)
        dtable = None
        if need_table:
            dtable = torch.zeros(tshape, dtype=torch.float32, device=g.device)
            if npos == 2:
                hashgrid_bwd_rays(pos[0], pos[1], d_enc, spec, box6, dtable)
            else:
                hashgrid_bwd(pos[0], d_enc, spec, box6, dtable)
        return (None, None, None, dtable, None, None, None, None, None, *grads)


def grad_node_of(t: Optional[Tensor]):
    """The autograd node that receives `t`'s gradient (AccumulateGrad for a leaf), or None.  Call with grad mode on,
    outside any autograd.Function."""
    if t is None or not t.requires_grad or not torch.is_grad_enabled():
        return None
    return t.grad_fn if t.grad_fn is not None else t.view_as(t).grad_fn.next_functions[0][0]


def grad_ctx(t: Optional[Tensor]):
    """(grad_node_of(t), is grad mode on) -- what ExpertFieldFn / RoutedFieldFn take as `table_node(s)`: grad mode is
    always off inside an autograd.Function's forward, and a forward that will never be differentiated need not write
    the encoding."""
    return grad_node_of(t), torch.is_grad_enabled()


def engine_wants(node) -> bool:
    """Inside a backward: will the running autograd pass deliver a gradient to `node`?  (True when unknown.)"""
    if node is None:
        return True
    try:
        return bool(torch._C._will_engine_execute_node(node))
    except Exception:
        return True


class BackgroundFn(torch.autograd.Function):
    """Background head (models/inr/meta_container.py:347-382): d (N,>=3) -> rgb (N,3); differentiable w.r.t. the four
    bg_mlp tensors (directions come from rays built under no_grad)."""

    @staticmethod
    def forward(ctx_, d, w1, b1, w2, b2, out_half):
        if not d.is_cuda:
            raise RuntimeError("background head needs CUDA tensors; there is no CPU path")
        if d.dtype != torch.float32 or d.stride(1) != 1:
            d = d.float().contiguous()                  # a (N,3) view of packed rays is read in place (row stride)
        ws = [dev_f32(w, "bg_mlp weight") for w in (w1, b1, w2, b2)]
        N, dev, Hb = d.shape[0], d.device, ws[0].shape[0]
        out = torch.empty(N, 3, dtype=torch.float16 if out_half else torch.float32, device=dev)
        check(lib().acn_background_fwd(ctx(dev), ptr(d), N, d.stride(0) if N else 3, *[ptr(w) for w in ws], Hb, ptr(out), _dt(out),
                                       stream(dev)))
        ctx_.save_for_backward(d, *ws)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx_, g):
        d, *ws = ctx_.saved_tensors
        N, dev, Hb = d.shape[0], d.device, ws[0].shape[0]
        need = ctx_.needs_input_grad[1:5]
        flat = torch.zeros(sum(w.numel() for w, n in zip(ws, need) if n), dtype=torch.float32, device=dev)
        grads, off = [], 0
        for w, n in zip(ws, need):
            grads.append(flat[off:off + w.numel()].view(w.shape) if n else None)
            off += w.numel() if n else 0
        g = g.contiguous().float()
        check(lib().acn_background_bwd(ctx(dev), ptr(d), N, d.stride(0) if N else 3, *[ptr(w) for w in ws], Hb, ptr(g),
                                       *[ptr(x) for x in grads], stream(dev)))
        return (None, *grads, None)


# ------------------------------------------------------------------------------------------ stage 4
def composite_fwd(rgb_sigma: Tensor, t: Tensor, bg: Optional[Tensor], sigma_scale: float):
    N, S = t.shape
    dev = t.device
    rgb = torch.empty(N, 3, dtype=torch.float32, device=dev)
    depth = torch.empty(N, dtype=torch.float32, device=dev)
    weights = torch.empty(N, S, dtype=torch.float32, device=dev)
    acc = torch.empty(N, dtype=torch.float32, device=dev)
    check(lib().acn_composite_fwd(ctx(dev), ptr(rgb_sigma), ptr(t), ptr(bg), N, S, float(sigma_scale), ptr(rgb), ptr(depth),
                                  ptr(weights), ptr(acc), stream(dev)))
    return rgb, depth, weights, acc


class CompositeFn(torch.autograd.Function):
    """nerfs/ray_rendering.py:114-165 volume_render (raw_rgb = raw_sigma = False)."""

    @staticmethod
    def forward(ctx_, rgb_sigma, t, bg, sigma_scale):
        rs = dev_f32(rgb_sigma, "rgb_sigma")
        t = dev_f32(t, "t_vals")
        bgc = dev_f32(bg.to(rs.device), "bg_rgb") if bg is not None else None
        out = composite_fwd(rs, t, bgc, sigma_scale)
        ctx_.save_for_backward(rs, t, *(b for b in (bgc,) if b is not None))
        ctx_.sigma_scale, ctx_.has_bg = float(sigma_scale), bgc is not None
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx_, g_rgb, g_depth, g_weights, g_acc):
        saved = ctx_.saved_tensors
        rs, t = saved[0], saved[1]
        bg = saved[2] if ctx_.has_bg else None
        N, S = t.shape
        dev = t.device
        c = lambda g: None if g is None else g.contiguous().float()
        g_rgb, g_depth, g_weights, g_acc = c(g_rgb), c(g_depth), c(g_weights), c(g_acc)
        d = torch.empty(N, S, 4, dtype=torch.float32, device=dev)
        need_bg = ctx_.has_bg and ctx_.needs_input_grad[2]
        d_bg = torch.zeros(N, 3, dtype=torch.float32, device=dev) if need_bg else None
        check(lib().acn_composite_bwd(ctx(dev), ptr(rs), ptr(t), ptr(bg), N, S, ctx_.sigma_scale, ptr(g_rgb), ptr(g_depth),
                                      ptr(g_weights), ptr(g_acc), ptr(d), ptr(d_bg), stream(dev)))
        return d.view_as(rs), None, d_bg, None


# ------------------------------------------------------------------------------------------ loss epilogue
def color_mse(pred: Tensor, gt: Tensor, color_space: str, reduction: str = "mean", want_grad: bool = False):
    """nerfs/color_space.py:22-66 + F.mse_loss in one launch.  -> (loss | per-element squared errors, d loss / d pred | None)"""
    cs = str(color_space).lower()
    if cs not in _lib.COLOR_SPACE:
        raise ValueError(f"Invalid color_space={color_space!r}; use 'linear'|'srgb'|'identity'")
    if reduction not in ("mean", "sum", "none"):
        raise ValueError(f"{reduction} is not a valid value for reduction")
    p = dev_f32(pred, "pred_rgb")
    g = dev_f32(gt.to(p.device), "gt_rgb")
    if g.shape != p.shape:
        g = g.expand_as(p).contiguous()
    dev, n = p.device, p.numel()
    reduced = reduction != "none"
    loss = torch.empty((), dtype=torch.float32, device=dev) if reduced else None
    elem = torch.empty_like(p) if not reduced else None
    dpred = torch.empty_like(p) if want_grad else None
    partial = torch.empty(_lib.LOSS_PARTIALS, dtype=torch.float64, device=dev) if reduced else None
    check(lib().acn_color_mse(ctx(dev), ptr(p), ptr(g), n, _lib.COLOR_SPACE[cs], int(reduction == "mean"), ptr(loss),
                              ptr(elem), ptr(dpred), ptr(partial), stream(dev)))
    return (loss if reduced else elem), dpred


class ColorMSEFn(torch.autograd.Function):
    """color_space_transformer(pred, gt, cs) -> F.mse_loss (nerfs/losses.py:29-32); the ground truth carries no gradient."""

    @staticmethod
    def forward(ctx_, pred, gt, color_space, reduction):
        out, dpred = color_mse(pred, gt, color_space, reduction, want_grad=pred.requires_grad)
        ctx_.save_for_backward(dpred)
        ctx_.shape, ctx_.dtype = pred.shape, pred.dtype
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx_, g_out):
        (dpred,) = ctx_.saved_tensors
        return (dpred * g_out).view(ctx_.shape).to(ctx_.dtype), None, None, None


# ------------------------------------------------------------------------------------------ stage 5
def route_points(pts: Tensor, centroids: Tensor, dims: int, margin: float, want_counts: bool = False):
    """-> (weights (P,K) | None, hard (P,) int32 | None, counts (K,) int32 | None)"""
    if not pts.is_cuda:
        raise RuntimeError("route_points needs CUDA tensors; there is no CPU path")
    if pts.dtype != torch.float32 or pts.stride(-1) != 1:
        pts = pts.float().contiguous()
    cen = dev_f32(centroids.to(pts.device), "centroids")
    P, K = pts.shape[0], cen.shape[0]
    dev = pts.device
    soft = margin > 1.0
    w = torch.empty(P, K, dtype=torch.float32, device=dev) if soft else None
    h = torch.empty(P, dtype=torch.int32, device=dev) if not soft else None
    counts = torch.zeros(K, dtype=torch.int32, device=dev) if want_counts else None
    check(lib().acn_route_points(ctx(dev), ptr(pts), P, pts.stride(0) if P else 3, ptr(cen), K, dims, float(margin),
                                 ptr(w), ptr(h), ptr(counts), stream(dev)))
    return w, h, counts


def route_rays_voronoi(rays: Tensor, S: int, centroids: Tensor, dims: int, margin: float,
                       aabb_out: Optional[Tuple[Tensor, Tensor, Optional[Tensor]]] = None) -> Tensor:
    """Ray -> expert masks (N,K) bool.  aabb_out = (mins (K,3) fp32, maxs (K,3) fp32, counts (K,) int64 | None) on the device
    are UPDATED with the samples assigned to each expert (scripts/create_clusters.py:386-556 update_aabbs)."""
    rays = dev_f32(rays, "rays")
    cen = dev_f32(centroids.to(rays.device), "centroids")
    N, K = rays.shape[0], cen.shape[0]
    mask = torch.empty(N, K, dtype=torch.uint8, device=rays.device)
    mins = maxs = counts = None
    if aabb_out is not None:
        mins, maxs, counts = aabb_out
        assert mins.dtype == torch.float32 and maxs.dtype == torch.float32 and mins.shape == (K, 3) == maxs.shape
        assert mins.is_contiguous() and maxs.is_contiguous() and mins.device == rays.device == maxs.device
        assert counts is None or (counts.dtype == torch.int64 and counts.shape == (K,) and counts.device == rays.device)
    check(lib().acn_route_rays_voronoi(ctx(rays.device), ptr(rays), N, S, ptr(u_lin(S, rays.device)), ptr(cen), K, dims,
                                       float(margin), ptr(mask), ptr(mins), ptr(maxs), ptr(counts), stream(rays.device)))
    return mask.bool()


def bucket_points(id6: Tensor, weights: Optional[Tensor], hard: Optional[Tensor], K: int, offsets: Tensor, total: int):
    """-> sel (total,) int32, xd (total,6), w (total,)"""
    dev = id6.device
    P = id6.shape[0]
    sel = torch.empty(total, dtype=torch.int32, device=dev)
    xd = torch.empty(total, 6, dtype=torch.float32, device=dev)
    w = torch.empty(total, dtype=torch.float32, device=dev)
    cursor = torch.zeros(K, dtype=torch.int32, device=dev)
    check(lib().acn_bucket_points(ctx(dev), ptr(id6), P, ptr(weights), ptr(hard), K, ptr(offsets), ptr(cursor), ptr(sel),
                                  ptr(xd), ptr(w), stream(dev)))
    return sel, xd, w


def route_count_rays(rays: Tensor, t: Tensor, centroids: Tensor, dims: int, margin: float, want_support: bool = False,
                     ray_major=False):
    """Rows per expert for the samples o + d*t of `rays` -> (K,) int32 (device) [, support sets (N*S,) uint16]."""
    rays, t = dev_f32(rays, "rays"), dev_f32(t, "t_vals")
    cen = dev_f32(centroids.to(rays.device), "centroids")
    N, S = t.shape
    counts = torch.zeros(cen.shape[0], dtype=torch.int32, device=rays.device)
    support = torch.empty(N * S, dtype=torch.uint16, device=rays.device) if want_support else None
    check(lib().acn_route_count_rays(ctx(rays.device), ptr(rays), ptr(t), N, S, ptr(cen), cen.shape[0], dims, float(margin),
                                     *_ray_major_args(ray_major), ptr(support), ptr(counts), stream(rays.device)))
    return (counts, support) if want_support else counts


def bucket_plan(counts: Tensor, cap: int, overflow: Optional[Tensor] = None):
    """Device-side bucket layout from device-side counts (acn_bucket_plan): -> seg (K+1,) int32 exclusive offsets clamped
    to `cap` rows, limit (K,) rows per expert that fit, cursor (K,) zeros.  `overflow` (1,) int32 is set when rows were cut."""
    K, dev = counts.shape[0], counts.device
    seg = torch.empty(K + 1, dtype=torch.int32, device=dev)
    limit = torch.empty(K, dtype=torch.int32, device=dev)
    cursor = torch.empty(K, dtype=torch.int32, device=dev)
    check(lib().acn_bucket_plan(ctx(dev), ptr(counts), K, int(cap), ptr(seg), ptr(limit), ptr(cursor), ptr(overflow), stream(dev)))
    return seg, limit, cursor


def shard_plan(all_counts: Tensor, rank: int, cap_local: int, cap_peer: int, overflow: Optional[Tensor] = None):
    """acn_shard_plan: all_counts (world, K) int32 on the device -> seg_local (K+1,), limit (K,), row_off (K,), seg_recv
    (K/world+1,), cursor (K,) -- the whole exchange layout of a sharded routed step, decided on the device."""
    world, K = all_counts.shape
    dev = all_counts.device
    i32 = lambda n: torch.empty(n, dtype=torch.int32, device=dev)
    seg_local, limit, row_off, seg_recv, cursor = i32(K + 1), i32(K), i32(K), i32(K // world + 1), i32(K)
    check(lib().acn_shard_plan(ctx(dev), ptr(all_counts), world, K, int(rank), int(cap_local), int(cap_peer), ptr(seg_local), ptr(limit),
                               ptr(row_off), ptr(seg_recv), ptr(cursor), ptr(overflow), stream(dev)))
    return seg_local, limit, row_off, seg_recv, cursor


def route_bucket_rays(rays: Tensor, t: Tensor, centroids: Tensor, dims: int, margin: float, offsets: Tensor, total: int,
                      support: Optional[Tensor] = None, ray_major=False, row_base: Optional[Tensor] = None,
                      row_off: Optional[Tensor] = None, row_limit: Optional[Tensor] = None, cursor: Optional[Tensor] = None):
    """-> sel (total,) int32 sample index, xd (total,6) [xyz, dir] rows, w (total,) blend weights; expert k's rows lie in
    [offsets[k], offsets[k] + counts[k])."""
    rays, t = dev_f32(rays, "rays"), dev_f32(t, "t_vals")
    cen = dev_f32(centroids.to(rays.device), "centroids")
    dev, K = rays.device, cen.shape[0]
    N, S = t.shape
    sel = torch.empty(total, dtype=torch.int32, device=dev)
    xd = torch.empty(total, 6, dtype=torch.float32, device=dev) if row_base is None else None   # else: rows go to row_base[k]
    w = torch.empty(total, dtype=torch.float32, device=dev)
    if cursor is None:
        cursor = torch.zeros(K, dtype=torch.int32, device=dev)
    assert row_base is None or (row_base.dtype == torch.int64 and row_off.dtype == torch.int32 and row_base.numel() == K == row_off.numel())
    check(lib().acn_route_bucket_rays(ctx(dev), ptr(rays), ptr(t), N, S, ptr(cen), K, dims, float(margin), *_ray_major_args(ray_major),
                                      ptr(support), ptr(offsets), ptr(cursor), ptr(sel), ptr(xd), ptr(w), ptr(row_base), ptr(row_off),
                                      ptr(row_limit), stream(dev)))
    return sel, xd, w


def bin_indices(hard: Tensor, K: int, offsets: Tensor, total: int) -> Tensor:
    """Indices i grouped by hard[i] (int32, -1 = skip): group k occupies [offsets[k], offsets[k] + count_k) -> (total,) int32."""
    dev = hard.device
    sel = torch.empty(total, dtype=torch.int32, device=dev)
    if total == 0 or hard.shape[0] == 0:
        return sel
    cursor = torch.zeros(K, dtype=torch.int32, device=dev)
    check(lib().acn_bucket_points(ctx(dev), None, hard.shape[0], None, ptr(hard), K, ptr(offsets), ptr(cursor), ptr(sel), None,
                                  None, stream(dev)))
    return sel


def dispatch_points(id6: Tensor, weights: Optional[Tensor], hard: Optional[Tensor], K: int, offsets: Tensor, total: int,
                    row_base: Tensor, row_off: Tensor):
    """bucket_points whose [xyz,dir] rows are stored into per-expert destination buffers (peer memory): row_base (K,)
    int64 device addresses, row_off (K,) int32 first rows.  -> sel (total,) int32, w (total,) local."""
    dev = id6.device
    sel = torch.empty(total, dtype=torch.int32, device=dev)
    w = torch.empty(total, dtype=torch.float32, device=dev)
    cursor = torch.zeros(K, dtype=torch.int32, device=dev)
    assert row_base.dtype == torch.int64 and row_off.dtype == torch.int32 and row_base.numel() == K and row_off.numel() == K
    check(lib().acn_dispatch_points(ctx(dev), ptr(id6), id6.shape[0], ptr(weights), ptr(hard), K, ptr(offsets), ptr(cursor), ptr(sel),
                                    ptr(w), ptr(row_base), ptr(row_off), stream(dev)))
    return sel, w


class RoutedFieldFn(torch.autograd.Function):
    """Several experts on their buckets of ONE routed row list whose bucket sizes only the device knows
    (models/inr/meta_container.py:306-337 without its K host syncs): xd (cap,>=6) [xyz, dir] rows, seg (n+1,) int32 device
    offsets -- expert i of `experts` evaluates rows [seg[i], seg[i+1]) -- -> y (cap,4) [rgb, sigma] (rows outside the
    ranges are never written nor read).  Every kernel takes the range as a device pointer; nothing is read back.
    experts: list of (GridSpec, box6).  Flat tensor arguments: per expert its hash table, then its 14 MLP tensors.
    y_out: optional (cap,4) fp32 buffer to write into (a peer-mapped buffer in the sharded container)."""

    @staticmethod
    def forward(ctx_, xd, seg, half, experts, table_nodes, y_out, *tensors):
        n = len(experts)
        assert len(tensors) == 15 * n and seg.dtype == torch.int32 and seg.numel() == n + 1
        if xd.dtype != torch.float32 or not xd.is_contiguous():
            xd = xd.float().contiguous()
        cap, dev = xd.shape[0], xd.device
        tables = [dev_f32(tensors[15 * i], "hash_table") for i in range(n)]
        wss = [[dev_f32(w, "MLP weight") for w in tensors[15 * i + 1:15 * i + 15]] for i in range(n)]
        E = experts[0][0].L * experts[0][0].F
        fused = [fused_fwd_ok(spec, wss[i], half) for i, (spec, _) in enumerate(experts)]
        table_nodes, grad_on = table_nodes if isinstance(table_nodes, tuple) else (table_nodes, True)
        want_enc = grad_on and any(ctx_.needs_input_grad)   # inference with every expert fused never materialises the encoding
        enc = (torch.empty(cap, E, dtype=torch.float16 if half else torch.float32, device=dev) if (want_enc or not all(fused))
               else xd.new_empty(0))
        y = y_out if y_out is not None else torch.empty(cap, 4, dtype=torch.float32, device=dev)
        assert y.shape == (cap, 4) and y.dtype == torch.float32 and y.is_contiguous()
        dirs = xd[:, 3:]
        L_ = lib()
        for i, (spec, box6) in enumerate(experts):
            rng = seg[i:i + 2]
            if fused[i]:
                render_expert_fwd((xd,), tables[i], spec, box6, dirs, xd.stride(0), 1, wss[i], want_enc, rng=rng, enc=enc, out=y)
                continue
            _, H, G, C = _field_dims(wss[i])
            check(L_.acn_hashgrid_fwd(ctx(dev), ptr(xd), cap, xd.stride(0), ptr(rng), ptr(box6), ptr(tables[i]), spec.L, spec.F,
                                      spec.log2T, ptr(_grid_res(spec, dev)), spec.interp, ptr(enc), _dt(enc), None, stream(dev)))
            st = pack_weights(wss[i])
            check(L_.acn_field_fwd(ctx(dev), ptr(enc), _dt(enc), ptr(dirs), xd.stride(0), 1, cap, ptr(rng), E, H, G, C, C_.byref(st),
                                   F16 if half else F32, ptr(y), stream(dev)))
        ctx_.save_for_backward(enc, xd, seg, *[b for _, b in experts if b is not None], *[w for ws in wss for w in ws])
        ctx_.meta = ([s_ for s_, _ in experts], [b is not None for _, b in experts], half, [t.shape for t in tables])
        ctx_.table_nodes = table_nodes
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx_, g):
        specs, has_box, half, tshapes = ctx_.meta
        n = len(specs)
        saved = ctx_.saved_tensors
        enc, xd, seg = saved[0], saved[1], saved[2]
        nb = sum(has_box)
        boxes_it = iter(saved[3:3 + nb])
        boxes = [next(boxes_it) if hb else None for hb in has_box]
        wflat = saved[3 + nb:]
        g = g.contiguous().float()
        cap, dev = xd.shape[0], xd.device
        dirs = xd[:, 3:]
        L_ = lib()
        out = []
        d_enc = None
        for i in range(n):
            spec, box6, ws = specs[i], boxes[i], wflat[14 * i:14 * i + 14]
            E, H, G, C = _field_dims(ws)
            rng = seg[i:i + 2]
            need_table = ctx_.needs_input_grad[6 + 15 * i] and engine_wants(ctx_.table_nodes[i])
            need_w = ctx_.needs_input_grad[7 + 15 * i:21 + 15 * i]
            flat = torch.zeros(sum(w.numel() for w, nd in zip(ws, need_w) if nd), dtype=torch.float32, device=dev)
            grads, off = [], 0
            for w, nd in zip(ws, need_w):
                grads.append(flat[off:off + w.numel()].view(w.shape) if nd else None)
                off += w.numel() if nd else 0
            wst, gst = pack_weights(ws), pack_weights(grads)
            dtable = torch.zeros(tshapes[i], dtype=torch.float32, device=dev) if need_table else None
            if (FUSED_EXPERT_BWD and need_table and half and spec.F == 2 and spec.L in (8, 16) and spec.interp != 0):
                check(L_.acn_render_expert_bwd(ctx(dev), ptr(xd), xd.stride(0), None, None, cap, 1, ptr(rng), ptr(box6), spec.L, spec.F,
                                               spec.log2T, ptr(_grid_res(spec, dev)), spec.interp, ptr(enc), ptr(dirs), xd.stride(0), 1,
                                               H, G, C, C_.byref(wst), ptr(g), C_.byref(gst), ptr(dtable), stream(dev)))
            else:
                if need_table and d_enc is None:
                    d_enc = torch.empty(cap, E, dtype=torch.float32, device=dev)
                check(L_.acn_field_bwd(ctx(dev), ptr(enc), _dt(enc), ptr(dirs), xd.stride(0), 1, cap, ptr(rng), E, H, G, C, C_.byref(wst),
                                       F16 if half else F32, ptr(g), C_.byref(gst), ptr(d_enc) if need_table else None, F32, stream(dev)))
                if need_table:
                    check(L_.acn_hashgrid_bwd(ctx(dev), ptr(xd), cap, xd.stride(0), ptr(rng), ptr(box6), spec.L, spec.F, spec.log2T,
                                              ptr(_grid_res(spec, dev)), spec.interp, ptr(d_enc), F32, ptr(dtable), stream(dev)))
            out += [dtable, *grads]
        return (None, None, None, None, None, None, *out)


class BlendRangesFn(torch.autograd.Function):
    """out[sel[i]] += y[i] * w[i] for the rows of every expert's range, one launch per expert in expert order (the
    reference's summation order, meta_container.py:321 / :336), ranges read on the device.  seg (n+1,) int32."""

    @staticmethod
    def forward(ctx_, y, w, sel, seg, N):
        y = y.contiguous()
        dev = y.device
        out = torch.zeros(N, 4, dtype=torch.float32, device=dev)
        for i in range(seg.numel() - 1):
            check(lib().acn_blend_add(ctx(dev), ptr(y), ptr(w), ptr(sel), sel.shape[0], ptr(seg[i:i + 2]), None, ptr(out), stream(dev)))
        ctx_.save_for_backward(w, sel, seg)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx_, g):
        w, sel, seg = ctx_.saved_tensors
        g = g.contiguous()
        dev = g.device
        d_y = torch.empty(sel.shape[0], 4, dtype=torch.float32, device=dev)
        span = torch.stack([seg[0], seg[-1]])
        check(lib().acn_blend_bwd(ctx(dev), ptr(g), ptr(w), ptr(sel), sel.shape[0], ptr(span), None, ptr(d_y), stream(dev)))
        return d_y, None, None, None, None


class BlendFn(torch.autograd.Function):
    """out[sel[i]] += y[i] * w[i] (models/inr/meta_container.py:321 index_add_ / :336 index_copy_)
    for one expert's routed rows; `out` is threaded through so experts accumulate in k order."""

    @staticmethod
    def forward(ctx_, out, y, w, sel):
        y = y.contiguous()
        check(lib().acn_blend_add(ctx(y.device), ptr(y), ptr(w), ptr(sel), y.shape[0], None, None, ptr(out), stream(y.device)))
        ctx_.save_for_backward(w, sel)
        ctx_.mark_dirty(out)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx_, g):
        w, sel = ctx_.saved_tensors
        g = g.contiguous()
        d_y = torch.empty(sel.shape[0], 4, dtype=torch.float32, device=g.device)
        check(lib().acn_blend_bwd(ctx(g.device), ptr(g), ptr(w), ptr(sel), sel.shape[0], None, None, ptr(d_y), stream(g.device)))
        return g, d_y, None, None


# ------------------------------------------------------------------------------------------ cross-checks and probes
def hashgrid_bwd_plain(x: Tensor, dout: Tensor, spec: GridSpec, box6: Optional[Tensor], dtable: Tensor) -> None:
    """acn_hashgrid_bwd_plain: the per-(point, level) scatter, whatever the configuration (the run-length kernels'
    cross-check)."""
    if x.dtype != torch.float32 or x.stride(1) != 1:
        x = x.float().contiguous()
    dout = dout.contiguous()
    P, dev = x.shape[0], x.device
    check(lib().acn_hashgrid_bwd_plain(ctx(dev), ptr(x), P, x.stride(0) if P else 3, ptr(box6), spec.L, spec.F, spec.log2T,
                                       ptr(_grid_res(spec, dev)), spec.interp, ptr(dout), _dt(dout), ptr(dtable), stream(dev)))


# The functions below drive libacn_b200_debug.so (include/acn_b200_debug.h), not the product library.
def _dbg():
    return _lib.debug_lib(), _lib.debug_check, _lib.debug_ctx


def debug_umma_gemm(a: Tensor, w: Tensor) -> Tensor:
    """D = A @ W^T on one tcgen05 tile: A (128,K) fp16, W (N,K) fp16 -> (128,N) fp32."""
    assert a.dtype == torch.float16 and w.dtype == torch.float16 and a.shape[0] == 128
    l, chk, dctx = _dbg()
    a, w = a.contiguous(), w.contiguous()
    d = torch.empty(128, w.shape[0], dtype=torch.float32, device=a.device)
    chk(l.acn_debug_umma_gemm(dctx(a.device), ptr(a), ptr(w), w.shape[0], a.shape[1], ptr(d), stream(a.device)))
    return d


def debug_umma_gemm_ts(a: Tensor, w: Tensor) -> Tensor:
    """D = A @ W^T with A staged in tensor memory (tcgen05.st) and read by the MMA from there."""
    assert a.dtype == torch.float16 and w.dtype == torch.float16 and a.shape[0] == 128
    l, chk, dctx = _dbg()
    a, w = a.contiguous(), w.contiguous()
    d = torch.empty(128, w.shape[0], dtype=torch.float32, device=a.device)
    chk(l.acn_debug_umma_gemm_ts(dctx(a.device), ptr(a), ptr(w), w.shape[0], a.shape[1], ptr(d), stream(a.device)))
    return d


def debug_umma_rate(mode: int, M: int, N: int, nmma: int, reps: int = 200, issuers: int = 1) -> List[int]:
    """Average SM cycles for `nmma` back-to-back tcgen05.mma + commit + wait, per issuing thread (acn_debug_umma_rate)."""
    l, chk, dctx = _dbg()
    dev = torch.device("cuda", torch.cuda.current_device())
    out = torch.zeros(4, dtype=torch.int64, device=dev)
    chk(l.acn_debug_umma_rate(dctx(dev), mode, M, N, nmma, reps, issuers, ptr(out), stream(dev)))
    return out.cpu().tolist()[:issuers]


def debug_l2_probe(mode: int, bytes_per_access: int, buf: Tensor, iters: int, grid: int) -> int:
    """One launch of the L2 gather (mode 0) / RED (mode 1) micro-benchmark -> accesses issued (acn_debug_l2_probe)."""
    l, chk, dctx = _dbg()
    dev = buf.device
    sink = torch.zeros(4, dtype=torch.float32, device=dev)
    chk(l.acn_debug_l2_probe(dctx(dev), mode, bytes_per_access, ptr(buf), buf.numel() * buf.element_size(), iters, grid,
                             ptr(sink), stream(dev)))
    return grid * 256 * iters * 8


def debug_field_trace(buf: Optional[Tensor]) -> None:
    """Arm (int64 CUDA tensor of 1024 words) or disarm (None) the MLP kernels' timeline in the DEBUG library; drive the
    traced kernels with `use_debug_library()` active (tools/field_trace.py)."""
    l, chk, dctx = _dbg()
    dev = buf.device if buf is not None else torch.device("cuda", torch.cuda.current_device())
    if buf is not None:
        assert buf.dtype == torch.int64 and buf.numel() >= 1024 and buf.is_contiguous()
    chk(l.acn_debug_field_trace(dctx(dev), ptr(buf)))


def use_debug_library() -> None:
    """tools/ only: route every entry point of this module through libacn_b200_debug.so (same ABI, built with the kernel
    timelines compiled in)."""
    global lib, ctx, check
    lib, ctx, check = _lib.debug_lib, _lib.debug_ctx, _lib.debug_check
