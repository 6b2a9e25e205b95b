"""One process per GPU: the three ways the rendering hot path shards across a B200 box.

The reference is single-process (nerf_runner.py:52-56); its only collective is the AABB / count all-reduce of the
offline mask script (scripts/create_clusters.py:898-932).  Everything here is therefore new capability, laid over
`torch.distributed` (NCCL over NVLink 5 / NVSwitch on the GPUs, gloo in the CPU tests):

* expert sharding (`ExpertShardedContainer`): expert k lives on rank k % world.  Rays stay data parallel on their
  home rank (clip, sample, route, composite); routed samples travel to the owner of their expert(s) with ONE
  variable-size all-to-all each way -- 24 B/sample out ([xyz, dir]), 16 B/sample back ([rgb, sigma]) -- and the
  backward mirrors it (dL/d[rgb, sigma] out, nothing back: positions carry no gradient).  Overlap samples inside
  `boundary_margin` go to every owner with non-zero weight and are blended at home in expert order, exactly like
  models/inr/meta_container.py:306-337.
* per-expert meta-training (`active_module=cid`): experts share nothing but the background head and the global
  gradient clip (pipelines/offline_stage/meta_core.py:181-190) -> `sharded_clip_grad_norm_` (one scalar all-reduce).
* single-expert data parallel: `allreduce_grads_` sums the hash-table gradient (64 MiB at T = 2^19) and the 13 715
  MLP gradients (flattened into one message).

Only host logic lives here; all arithmetic on samples is in the CUDA kernels behind `ops`."""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from . import ops


# --------------------------------------------------------------------------------------------- exchange primitives
class _AllToAllRows(torch.autograd.Function):
    """Variable-size row exchange y = all_to_all(x) with autograd: the backward is the same exchange with the
    splits swapped.  x: (sum(send_splits), C) grouped by destination rank."""

    @staticmethod
    def forward(ctx, x: Tensor, send_splits: List[int], recv_splits: List[int], group):
        x = x.contiguous()
        y = x.new_empty((sum(recv_splits),) + tuple(x.shape[1:]))
        dist.all_to_all_single(y, x, output_split_sizes=recv_splits, input_split_sizes=send_splits, group=group)
        ctx.splits, ctx.group = (send_splits, recv_splits), group
        return y

    @staticmethod
    def backward(ctx, g: Tensor):
        send_splits, recv_splits = ctx.splits
        g = g.contiguous()
        gx = g.new_empty((sum(send_splits),) + tuple(g.shape[1:]))
        dist.all_to_all_single(gx, g, output_split_sizes=send_splits, input_split_sizes=recv_splits, group=ctx.group)
        return gx, None, None, None


def all_to_all_rows(x: Tensor, send_splits: Sequence[int], recv_splits: Sequence[int], group=None) -> Tensor:
    return _AllToAllRows.apply(x, [int(v) for v in send_splits], [int(v) for v in recv_splits], group)


def expert_owner_layout(K: int, world: int) -> int:
    """Experts per rank m.  Ownership is INTERLEAVED: expert k lives on rank k % world as its local expert k // world,
    so neighbouring Voronoi cells -- the ones a single camera view sees -- sit on different GPUs."""
    if K % world != 0:
        raise ValueError(f"{K} experts do not shard evenly over {world} ranks")
    return K // world


def expert_send_order(K: int, world: int) -> List[int]:
    """Global expert ids in the order their rows must be laid out for the exchange: destination rank major, local
    expert minor."""
    m = expert_owner_layout(K, world)
    return [e * world + r for r in range(world) for e in range(m)]


def gather_counts(counts_dev: Tensor, group=None) -> Tensor:
    """counts (K,) on the device -> (world, K) host matrix of every rank's per-expert row counts: one collective and
    ONE device-to-host copy give both the send and the receive splits of a routed step."""
    world = dist.get_world_size(group)
    c = counts_dev.to(torch.int64).contiguous()
    parts = [torch.empty_like(c) for _ in range(world)]
    dist.all_gather(parts, c, group=group)
    return torch.stack(parts).cpu()


def exchange_segments(all_counts: Tensor, rank: int) -> Tuple[List[int], List[int], List[Tuple[int, int, int]]]:
    """all_counts (world, K) -> (send_splits, recv_splits, segments) for `rank`.  The receive buffer is ordered by
    source rank, then local expert; segments lists (local expert, start, length) in buffer order."""
    world, K = all_counts.shape
    m = expert_owner_layout(K, world)
    ac = all_counts.tolist()
    send_splits = [int(sum(ac[rank][e * world + r] for e in range(m))) for r in range(world)]
    recv_splits, segs, pos = [], [], 0
    for src in range(world):
        tot = 0
        for e in range(m):
            n = int(ac[src][e * world + rank])
            segs.append((e, pos, n))
            pos += n
            tot += n
        recv_splits.append(tot)
    return send_splits, recv_splits, segs


def routed_exchange(xd: Tensor, all_counts: Tensor, local_fields, group=None) -> Tensor:
    """The sample round trip of one routed step.

    xd (total, C>=6): this rank's routed rows laid out in `expert_send_order`; all_counts (world, K) host matrix from
    `gather_counts`; local_fields: m callables, local_fields[e](rows (M,C)) -> (M,4), the experts this rank owns.
    -> y (total, 4) in the order of `xd`, differentiable w.r.t. whatever the owners' fields depend on (their
    gradients arrive through the backward all-to-all).  Every (source rank, expert) segment of the receive buffer is
    evaluated where it lies -- no gather / scatter copies on either pass."""
    rank = dist.get_rank(group)
    send_splits, recv_splits, segs = exchange_segments(all_counts, rank)
    assert len(local_fields) == expert_owner_layout(all_counts.shape[1], all_counts.shape[0])
    with torch.no_grad():                                   # positions / directions carry no gradient
        rows = all_to_all_rows(xd, send_splits, recv_splits, group)
    pieces = [local_fields[e](rows[s:s + n]).float() if n else rows.new_zeros((0, 4), dtype=torch.float32) for e, s, n in segs]
    y_recv = pieces[0] if len(pieces) == 1 else torch.cat(pieces)
    # The return trip's backward is a collective: every rank must run it, also one that received nothing or owns no
    # trainable parameter -- give autograd a reason to.
    if torch.is_grad_enabled() and not y_recv.requires_grad:
        y_recv = y_recv.detach().requires_grad_()
    return all_to_all_rows(y_recv, recv_splits, send_splits, group)


# --------------------------------------------------------------------------------------------- peer-memory exchange
class PeerExchange:
    """Symmetric (peer-mapped) buffers for the routed exchange of one rank group: every rank allocates the same three
    buffers -- rows (cap,6), y (cap,4), dy (cap,4) fp32 -- and maps everybody else's over NVLink
    (torch.distributed._symmetric_memory).  The dispatch kernel stores routed rows straight into the owners' `rows`;
    the blend kernel loads the owners' `y` straight out of peer memory; the blend backward stores dL/dy straight into
    the owners' `dy`.  Cross-GPU ordering is a device-side barrier on the stream (signal pads), no host involvement."""

    def __init__(self, cap_rows: int, device: torch.device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.cap = int(cap_rows)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.rows = symm.empty((self.cap, 6), dtype=torch.float32, device=device)
        self.y = symm.empty((self.cap, 4), dtype=torch.float32, device=device)
        self.dy = symm.empty((self.cap, 4), dtype=torch.float32, device=device)
        self.h_rows = symm.rendezvous(self.rows, self.group)
        self.h_y = symm.rendezvous(self.y, self.group)
        self.h_dy = symm.rendezvous(self.dy, self.group)
        self.rows_ptrs = [int(p) for p in self.h_rows.buffer_ptrs]

    def peer_y(self, r: int) -> Tensor:
        return self.y if r == self.rank else self.h_y.get_buffer(r, (self.cap, 4), torch.float32)

    def peer_dy(self, r: int) -> Tensor:
        return self.dy if r == self.rank else self.h_dy.get_buffer(r, (self.cap, 4), torch.float32)


class _PeerCombine(torch.autograd.Function):
    """out[sel] += w * y over all experts, y read from the owners' symmetric buffers; backward stores w * dL/dout[sel]
    into the owners' dy buffers and hands the owner ITS dy as the gradient of what it put into y."""

    @staticmethod
    def forward(ctx, y_local: Tensor, px: "PeerExchange", M: int, plan, sel: Tensor, wsel: Tensor, N: int):
        # plan: list of (k, owner, local_off, remote_off, n) in expert order
        if M:
            px.y[:M].copy_(y_local)
        px.h_y.barrier()                                     # every owner's y is complete and visible
        out = torch.zeros(N, 4, dtype=torch.float32, device=sel.device)
        for k, owner, lo, ro, n in plan:                     # blend in expert order, like the reference
            if n:
                yk = px.peer_y(owner)[ro:ro + n]
                ops.check(ops.lib().acn_blend_add(ops.ctx(out.device), ops.ptr(yk), ops.ptr(wsel[lo:lo + n]), ops.ptr(sel[lo:lo + n]), n,
                                                  None, None, ops.ptr(out), ops.stream(out.device)))
        ctx.px, ctx.M, ctx.plan = px, M, plan
        ctx.save_for_backward(sel, wsel)
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        px, M, plan = ctx.px, ctx.M, ctx.plan
        sel, wsel = ctx.saved_tensors
        g = g.contiguous()
        px.h_dy.barrier()                                    # nobody is still reading dy from the previous use
        for k, owner, lo, ro, n in plan:
            if n:
                dyk = px.peer_dy(owner)[ro:ro + n]
                ops.check(ops.lib().acn_blend_bwd(ops.ctx(g.device), ops.ptr(g), ops.ptr(wsel[lo:lo + n]), ops.ptr(sel[lo:lo + n]), n,
                                                  None, None, ops.ptr(dyk), ops.stream(g.device)))
        px.h_dy.barrier()                                    # every home rank's dL/dy has landed in my dy
        return px.dy[:M].clone(), None, None, None, None, None, None


class _PeerCombineRanges(torch.autograd.Function):
    """`_PeerCombine` without host-side sizes: out[sel] += w * y over all experts, every expert's rows read from its
    owner's symmetric y buffer at the device-side offsets of `acn_shard_plan` (local range seg_local[k:k+2], owner rows
    starting at row_off[k]); the backward stores w * dL/dout[sel] into the owners' dy buffers the same way and hands this
    rank ITS dy buffer as the gradient of what it put into y.  Blending stays in expert order."""

    @staticmethod
    def forward(ctx, y_local: Tensor, px: "PeerExchange", K: int, world: int, seg_local: Tensor, row_off: Tensor, sel: Tensor,
                wsel: Tensor, N: int):
        px.h_y.barrier()                                     # every owner's y is complete and visible
        dev = sel.device
        out = torch.zeros(N, 4, dtype=torch.float32, device=dev)
        L_, c, st = ops.lib(), ops.ctx(dev), ops.stream(dev)
        for k in range(K):                                   # expert order, like the reference (meta_container.py:306-337)
            ops.check(L_.acn_blend_add(c, ops.ptr(px.peer_y(k % world)), ops.ptr(wsel), ops.ptr(sel), sel.shape[0],
                                       ops.ptr(seg_local[k:k + 2]), ops.ptr(row_off[k:k + 1]), ops.ptr(out), st))
        ctx.px, ctx.K, ctx.world = px, K, world
        ctx.save_for_backward(sel, wsel, seg_local, row_off)
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        px, K, world = ctx.px, ctx.K, ctx.world
        sel, wsel, seg_local, row_off = ctx.saved_tensors
        g = g.contiguous()
        dev = g.device
        L_, c, st = ops.lib(), ops.ctx(dev), ops.stream(dev)
        px.h_dy.barrier()                                    # nobody is still reading dy from the previous use
        for k in range(K):
            ops.check(L_.acn_blend_bwd(c, ops.ptr(g), ops.ptr(wsel), ops.ptr(sel), sel.shape[0], ops.ptr(seg_local[k:k + 2]),
                                       ops.ptr(row_off[k:k + 1]), ops.ptr(px.peer_dy(k % world)), st))
        px.h_dy.barrier()                                    # every home rank's dL/dy has landed in my dy
        return px.dy, None, None, None, None, None, None, None, None


# --------------------------------------------------------------------------------------------- expert sharding
class ExpertShardedContainer(torch.nn.Module):
    """A `MetaContainer` whose experts live on different ranks.  Built from a full container description; every
    rank constructs the same object and then drops the experts it does not own (`shard_()`), so checkpoints keep the
    reference's `submodules.{k}.*` keys on the owner.

        forward(x (N,>=6)) -> (N,4)          the routed, blended field (meta_container.py:275-343)
        render via nerfs.ray_rendering.render_rays(model, rays, ..., active_module=None) as usual
    """

    def __init__(self, container, group=None, peer_rows: int = 0):
        """peer_rows > 0 switches the sample exchange from NCCL all-to-all to kernels that store to / load from peer
        memory directly (`PeerExchange` with that many rows of capacity per rank).  The render path (`forward_rays`)
        then runs WITHOUT any host read: the per-expert counts of all ranks are all-gathered on the device,
        `acn_shard_plan` lays the exchange out there, and the owners' kernels take device-side row ranges; rows that do
        not fit the capacity are dropped and `check_route_overflow()` reports it."""
        super().__init__()
        self.inner = container
        self.group = group
        self._peer_rows = int(peer_rows)
        self._px = None
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.K = len(container.submodules)
        self.m = expert_owner_layout(self.K, self.world)
        self.local_ids = [e * self.world + self.rank for e in range(self.m)]          # interleaved ownership
        self.use_bg_nerf = container.use_bg_nerf
        self.use_occ = False

    def shard_(self):
        """Free the experts this rank does not own (their slots stay in the ModuleList as empty modules)."""
        for k in range(self.K):
            if k not in self.local_ids:
                self.inner.submodules[k] = torch.nn.Identity()
        return self

    @property
    def submodules(self):
        return self.inner.submodules

    def background_color(self, d: Tensor) -> Tensor:
        return self.inner.background_color(d)

    def local_parameters(self) -> List[torch.nn.Parameter]:
        return [p for k in self.local_ids for p in self.inner.submodules[k].parameters()]

    def shared_parameters(self) -> List[torch.nn.Parameter]:
        return [p for n, p in self.inner.named_parameters() if not n.startswith("submodules.")]

    def forward(self, x: Tensor, params=None, active_module: Optional[int] = None) -> Tensor:
        if active_module is not None:
            if active_module not in self.local_ids:
                raise RuntimeError(f"expert {active_module} is owned by rank {active_module % self.world}, not {self.rank}")
            return self.inner(x, params=params, active_module=active_module)
        c = self.inner
        id6 = ops.dev_f32(x[:, :6], "points")
        with torch.no_grad():
            w, hard, counts = ops.route_points(id6, c.centroids, 2 if c.cluster_2d else 3, c.boundary_margin, want_counts=True)
        return self._exchange_and_blend(
            x.shape[0], counts, params,
            bucket=lambda offsets, total: ops.bucket_points(id6, w, hard, self.K, offsets, total),
            dispatch=lambda offsets, total, row_base, row_off: ops.dispatch_points(id6, w, hard, self.K, offsets, total, row_base, row_off))

    def forward_rays(self, rays: Tensor, t: Tensor, params=None, ray_major=False) -> Tensor:
        """rays (N,8), t (N,S) -> (N,S,4): the routed, blended field at the samples o + d*t, routed and bucketed (or
        dispatched into the owners' peer buffers) straight from the rays -- no (N*S,6) points, no (N*S,K) weights.
        `ray_major`: see MetaContainer.forward_rays."""
        c = self.inner
        if self.K > c.FUSED_ROUTE_MAX_EXPERTS:
            return self.forward(ops.points(rays, t), params=params).view(t.shape[0], t.shape[1], -1)
        N, S = t.shape
        dims = 2 if c.cluster_2d else 3
        with torch.no_grad():
            counts, support = ops.route_count_rays(rays, t, c.centroids, dims, c.boundary_margin, want_support=True, ray_major=ray_major)
        kw = dict(support=support, ray_major=ray_major)

        def bucket(offsets, total):
            return ops.route_bucket_rays(rays, t, c.centroids, dims, c.boundary_margin, offsets, total, **kw)

        def dispatch(offsets, total, row_base, row_off):
            sel, _, wsel = ops.route_bucket_rays(rays, t, c.centroids, dims, c.boundary_margin, offsets, total,
                                                 row_base=row_base, row_off=row_off, **kw)
            return sel, wsel

        if self._peer_rows > 0:
            return self._forward_rays_peer_nosync(rays, t, counts, support, ray_major, params).view(N, S, -1)
        return self._exchange_and_blend(N * S, counts, params, bucket, dispatch).view(N, S, -1)

    #: rows the local bucket arrays (sel, w) are allocated for, as a multiple of this rank's samples (soft routing)
    route_capacity_factor = 2.0

    def check_route_overflow(self) -> None:
        """Raises if a routed step since the last check needed more rows than the local arrays (`route_capacity_factor`)
        or an owner's peer buffer (`peer_rows`) hold.  Reads one device word: call it outside the timed steps."""
        f = getattr(self, "_overflow", None)
        if f is not None and int(f.item()) != 0:
            f.zero_()
            raise RuntimeError(f"sharded routed buckets overflowed (peer_rows={self._peer_rows}, "
                               f"route_capacity_factor={self.route_capacity_factor}); the affected steps dropped samples")

    def _forward_rays_peer_nosync(self, rays: Tensor, t: Tensor, counts: Tensor, support: Tensor, ray_major, params) -> Tensor:
        """The routed step over peer memory with everything decided on the device (no .cpu(), no .tolist()):
        all-gather of the counts -> acn_shard_plan -> dispatch straight into the owners' buffers -> barrier -> the owners'
        experts on device-side row ranges of their buffer, writing y into their symmetric buffer -> barrier -> blend from
        the peers' y in expert order.  The backward mirrors it through `_PeerCombineRanges` and `RoutedFieldFn`."""
        c = self.inner
        dev = rays.device
        N, S = t.shape
        P = N * S
        W, K = self.world, self.K
        if self._px is None:
            self._px = PeerExchange(self._peer_rows, dev, self.group)
        if getattr(self, "_row_base", None) is None:
            self._row_base = torch.tensor([self._px.rows_ptrs[k % W] for k in range(K)], dtype=torch.int64, device=dev)
            self._overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        px = self._px
        cap_local = P if c.boundary_margin <= 1.0 else int(min(float(K), float(self.route_capacity_factor)) * P)
        with torch.no_grad():
            all_counts = torch.empty(W, K, dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(all_counts, counts.contiguous(), group=self.group)       # stays on the device
            seg_local, limit, row_off, seg_recv, cursor = ops.shard_plan(all_counts, self.rank, cap_local, px.cap, self._overflow)
            px.h_rows.barrier()                               # the owners are done with the previous contents of `rows`
            sel, _, wsel = ops.route_bucket_rays(rays, t, c.centroids, 2 if c.cluster_2d else 3, c.boundary_margin, seg_local,
                                                 cap_local, support=support, ray_major=ray_major, row_base=self._row_base,
                                                 row_off=row_off, row_limit=limit, cursor=cursor)
            px.h_rows.barrier()                               # every rank's rows have landed in every owner
        # the owners' experts read their rows where they landed and write y where the peers will read it
        y_local = c._evaluate_segments(px.rows, seg_recv, self.local_ids, c._sub_params(params), y_out=px.y)
        if torch.is_grad_enabled() and not y_local.requires_grad:
            y_local = y_local.detach().requires_grad_()       # the combine's backward is collective: every rank must run it
        return _PeerCombineRanges.apply(y_local, px, K, W, seg_local, row_off, sel, wsel, P)

    def _exchange_and_blend(self, N: int, counts: Tensor, params, bucket, dispatch) -> Tensor:
        """counts (K,) rows per expert on this rank; bucket(offsets, total) -> (sel, xd, w) builds the local buckets,
        dispatch(offsets, total, row_base, row_off) -> (sel, w) stores the rows into the owners' peer buffers."""
        c = self.inner
        dev = counts.device
        sub_params = c._sub_params(params)
        with torch.no_grad():
            all_counts = gather_counts(counts, self.group)              # the step's one host read
            cnt = all_counts[self.rank]
            offsets = torch.zeros(self.K, dtype=torch.int32)
            pos = 0
            for k in expert_send_order(self.K, self.world):             # rows grouped by destination rank, then expert
                offsets[k] = pos
                pos += int(cnt[k])
            use_peer = self._peer_rows > 0 and int(all_counts.sum(0).reshape(self.m, self.world).sum(0).max()) <= self._peer_rows
            if not use_peer:
                sel, xd, wsel = bucket(offsets.to(dev), pos)
        fields = [(lambda rows, k=k: c.submodules[k](rows, params=sub_params[k])) for k in self.local_ids]
        if use_peer:
            return self._forward_peer(dispatch, dev, all_counts, cnt, offsets, pos, fields, N)
        y = routed_exchange(xd, all_counts, fields, self.group)
        # ties y into the graph even when this rank blends nothing (the exchange's backward is a collective) WITHOUT
        # touching values: the sum over zero rows is an exact 0 whatever y holds (0 * inf = NaN would poison every ray)
        out = torch.zeros(N, 4, dtype=torch.float32, device=dev) + y[:0].sum()
        off = offsets.tolist()
        for k in range(self.K):                               # blend in expert order, like the reference
            n = int(cnt[k])
            if n:
                sl = slice(off[k], off[k] + n)
                out = ops.BlendFn.apply(out, y[sl], wsel[sl], sel[sl])
        return out

    def _forward_peer(self, dispatch, dev, all_counts, cnt, offsets, total, fields, N):
        """The routed step over peer memory: dispatch kernel -> barrier -> owners' fields -> barrier -> blend from peers."""
        if self._px is None:
            self._px = PeerExchange(self._peer_rows, dev, self.group)
        px, W, m = self._px, self.world, self.m
        ac = all_counts.tolist()
        # first row of (source rank src, expert k) in the receive buffer of k's owner: buffer order is (src, local expert)
        def remote_off(src, k):
            owner, e = k % W, k // W
            return sum(ac[s][ee * W + owner] for s in range(src) for ee in range(m)) + sum(ac[src][ee * W + owner] for ee in range(e))
        plan = [(k, k % W, int(offsets[k]), remote_off(self.rank, k), int(cnt[k])) for k in range(self.K)]
        with torch.no_grad():
            row_base = torch.tensor([px.rows_ptrs[k % W] for k in range(self.K)], dtype=torch.int64, device=dev)
            row_off = torch.tensor([p[3] for p in plan], dtype=torch.int32, device=dev)
            px.h_rows.barrier()                               # the owners are done with the previous contents of `rows`
            sel, wsel = dispatch(offsets.to(dev), total, row_base, row_off)
            px.h_rows.barrier()                               # every rank's rows have landed in every owner
            _, recv_splits, segs = exchange_segments(all_counts, self.rank)
            M = sum(recv_splits)
            rows = px.rows[:M].clone()                        # the fields keep their input for the backward; free the buffer
        pieces = [fields[e](rows[s:s + n]).float() if n else rows.new_zeros((0, 4)) for e, s, n in segs]
        y_local = pieces[0] if len(pieces) == 1 else torch.cat(pieces)
        if torch.is_grad_enabled() and not y_local.requires_grad:
            y_local = y_local.detach().requires_grad_()       # the combine's backward is collective: every rank must run it
        return _PeerCombine.apply(y_local, px, M, plan, sel, wsel, N)


# --------------------------------------------------------------------------------------------- gradient plumbing
def allreduce_grads_(params: Iterable[torch.nn.Parameter], group=None, average: bool = True, flat_below: int = 1 << 20) -> None:
    """Sum (or average) .grad over the ranks: tensors below `flat_below` elements travel as ONE flattened message
    (the 14 MLP tensors of an expert are 13 715 floats), large ones (the hash table) on their own."""
    world = dist.get_world_size(group)
    small, big = [], []
    for p in params:
        if p.grad is None:
            continue
        (small if p.grad.numel() < flat_below else big).append(p.grad)
    for g in big:
        dist.all_reduce(g, group=group)
    if small:
        flat = torch.cat([g.reshape(-1) for g in small])
        dist.all_reduce(flat, group=group)
        off = 0
        for g in small:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
    if average and world > 1:
        for g in big + small:
            g.div_(world)


def sharded_clip_grad_norm_(owned: Iterable[torch.nn.Parameter], shared: Iterable[torch.nn.Parameter], max_norm: float,
                            group=None) -> Tensor:
    """torch.nn.utils.clip_grad_norm_ over ALL experts' parameters when each rank only holds its own
    (pipelines/offline_stage/meta_core.py:181-190 clips one global norm): `owned` gradients exist on exactly one rank,
    `shared` ones (background head) are replicated and already all-reduced.  One scalar all-reduce."""
    owned = [p for p in owned if p.grad is not None]
    shared = [p for p in shared if p.grad is not None]
    if owned or shared:
        dev = (owned + shared)[0].grad.device
    else:   # a rank without gradients still takes part in the all-reduce: the scalar must live where the backend expects it
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    sq = torch.zeros((), dtype=torch.float32, device=dev)
    for p in owned:
        sq = sq + p.grad.float().pow(2).sum()
    dist.all_reduce(sq, group=group)
    for p in shared:
        sq = sq + p.grad.float().pow(2).sum()
    total = sq.sqrt()
    coef = (max_norm / (total + 1e-6)).clamp(max=1.0)
    for p in owned + shared:
        p.grad.mul_(coef.to(p.grad.dtype))
    return total


def reduce_expert_aabbs(mins: Tensor, maxs: Tensor, counts: Tensor, group=None) -> Tuple[Tensor, Tensor, Tensor]:
    """Combine the per-expert sample AABBs / counts of rank-strided mask generation
    (scripts/create_clusters.py:898-904, 928-932): MIN / MAX / SUM all-reduces of (K,3), (K,3), (K,)."""
    dist.all_reduce(mins, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(maxs, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return mins, maxs, counts
