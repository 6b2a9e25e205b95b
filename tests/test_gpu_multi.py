"""Two-GPU parity of the expert-sharded container (adaptive_city_nerf_b200/distributed.py): every rank renders its
own rays through experts that live on different GPUs (NCCL all-to-all of routed samples, both directions, forward
and backward) and must reproduce what the single-process MetaContainer computes for the same rays and weights.
Needs >= 2 visible GPUs (run with `gpurun --gpus 2`); skipped otherwise."""
import os
import socket

import numpy as np
import pytest
import torch

import synth
from helpers import make_container

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rays(rank, dev, n=3000):
    o, d = synth.random_rays_in_box(50 + rank, n)
    from adaptive_city_nerf_b200.nerfs.scene_box import SceneBox
    from adaptive_city_nerf_b200.nerfs.ray_sampling import clamp_rays_near_far
    box = SceneBox(torch.from_numpy(synth.AABB_GLOBAL).to(dev))
    o, d = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
    tmin, tmax = box.ray_aabb_intersect(o, d)
    rays = torch.cat([o, d, tmin[:, None], tmax[:, None]], dim=1)
    rays, valid = clamp_rays_near_far(rays, (None, None))
    return rays[valid].contiguous()


def _worker(rank, world, port, margin, peer_rows, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from adaptive_city_nerf_b200.distributed import ExpertShardedContainer, allreduce_grads_, sharded_clip_grad_norm_
        from adaptive_city_nerf_b200.nerfs.ray_rendering import render_rays
        K, S = 4, 24
        full = make_container(K, synth.CENTROIDS_G22, synth.EXPERT_BOXES_G22, margin, True, seed0=300, device=dev).eval()
        # reference: all ranks' rays through the unsharded container, on this GPU
        ref_rgb, tot = [], 0.0
        for r in range(world):
            rr = _rays(r, dev)
            rgb, dep, _, _ = render_rays(full, rr, ray_samples=S, active_module=None)
            ref_rgb.append(rgb.detach())
            tot = tot + (rgb * torch.linspace(0.5, 1.5, rgb.numel(), device=dev).view_as(rgb)).sum() + dep.sum()
        tot.backward()
        ref_grads = {n: p.grad.clone() for n, p in full.named_parameters() if p.grad is not None}
        full.zero_grad(set_to_none=True)
        # sharded: this rank's rays only
        model = ExpertShardedContainer(full, peer_rows=peer_rows).shard_().eval()
        mine = _rays(rank, dev)
        rgb, dep, _, _ = render_rays(model, mine, ray_samples=S, active_module=None)
        err = float((rgb.detach() - ref_rgb[rank]).abs().max())
        loss = (rgb * torch.linspace(0.5, 1.5, rgb.numel(), device=dev).view_as(rgb)).sum() + dep.sum()
        loss.backward()
        allreduce_grads_(model.shared_parameters(), average=False)          # background head: replicated
        gerr = 0.0
        for n, p in full.named_parameters():
            if p.grad is None:
                continue
            a, b = p.grad.double(), ref_grads[n].double()
            gerr = max(gerr, float((a - b).norm() / (b.norm() + 1e-30)))
        n_owned = sum(1 for k in model.local_ids for _ in model.submodules[k].parameters())
        total = sharded_clip_grad_norm_(model.local_parameters(), model.shared_parameters(), 1e9)
        ref_total = float(torch.sqrt(sum(g.double().pow(2).sum() for g in ref_grads.values())))
        # the (N,>=6) point interface of the reference and the frame mode of the ray interface give the same field
        from adaptive_city_nerf_b200 import ops
        with torch.no_grad():
            t = ops.sample_stratified(mine, S, None)
            y_pts = model(ops.points(mine, t)).view(-1, S, 4)
            y_rays = model.forward_rays(mine, t, ray_major=True)
        perr = float((y_pts - y_rays).abs().max())
        overflowed = False
        try:
            model.check_route_overflow()
        except RuntimeError:
            overflowed = True
        ret[rank] = (err, gerr, n_owned, float(total), ref_total, perr, overflowed)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("margin,peer_rows", [(1.05, 0), (1.0, 0), (1.05, 1 << 18), (1.0, 1 << 18), (1.05, 1000)])
def test_expert_sharded_container_matches_single_process(margin, peer_rows):
    """peer_rows = 0: NCCL all-to-all exchange (one host read per step); 2^18: kernels storing to / loading from peer memory
    over NVLink (symmetric memory) with the whole exchange laid out on the device (acn_shard_plan: no host read);
    1000: a capacity the step overflows -- the surplus rows are dropped on the device and check_route_overflow() reports it."""
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), margin, peer_rows, ret), nprocs=2, join=True)
    for rank in range(2):
        err, gerr, n_owned, total, ref_total, perr, overflowed = ret[rank]
        if peer_rows == 1000:
            assert overflowed and np.isfinite(err)
            continue
        assert not overflowed
        assert perr == 0.0, (rank, perr)                    # points path == rays path (ray-major buckets), bit for bit
        assert err < 1e-5, (rank, err)                      # same kernels, same order of blending
        assert gerr < 1e-4, (rank, gerr)                    # float atomics reorder sums; nothing else differs
        assert n_owned == 2 * 15                            # two experts' 14 MLP tensors + table each (experts r and r + 2)
        assert abs(total - ref_total) / ref_total < 1e-4    # the sharded global norm is the global norm


def _worker_modes(rank, world, port, ret):
    """SURVEY 8e rows 1 and 4 on two GPUs: per-expert meta-training (each rank trains ITS expert; the reference clips ONE
    global gradient norm over all experts, pipelines/offline_stage/meta_core.py:181-190) and rank-strided Voronoi mask
    generation with the box / count reduction of scripts/create_clusters.py:928-932."""
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from adaptive_city_nerf_b200.data.cluster_masks import finalize_expert_boxes, new_expert_boxes, voronoi_masks_and_boxes
        from adaptive_city_nerf_b200.distributed import reduce_expert_aabbs
        from adaptive_city_nerf_b200.nerfs.ray_rendering import render_rays
        from adaptive_city_nerf_b200.optim import FusedAdam
        K, S = 2, 24
        cen = synth.CENTROIDS_G22[[0, 3]]
        boxes = synth.EXPERT_BOXES_G22[[0, 3]]

        def fresh():
            return make_container(K, cen, boxes, 1.0, False, seed0=400, device=dev).train()

        def expert_loss(m, k):
            rr = _rays(10 + k, dev, n=2000)
            jit = torch.rand(rr.shape[0], S, device=dev, generator=torch.Generator(device=dev).manual_seed(k))
            with torch.autocast("cuda", dtype=torch.float16):
                rgb, *_ = render_rays(m, rr, ray_samples=S, active_module=k, jitter=jit)
            return ((rgb - 0.25 * (k + 1)) ** 2).mean() * 50.0          # large enough for the clip to bite

        # one process, both experts, one optimizer, one global clip: what the reference's meta update does
        ref = fresh()
        opt = FusedAdam(ref.parameters(), lr=1e-2, eps=1e-15)
        for _ in range(3):
            opt.zero_grad(set_to_none=True)
            (expert_loss(ref, 0) + expert_loss(ref, 1)).backward()
            opt.step(max_norm=1.0)
        ref_norm = float(opt.last_norm)
        # sharded: this rank only ever touches expert `rank`; the clip still sees the global norm
        mine = fresh()
        own = list(mine.submodules[rank].parameters())
        opt2 = FusedAdam(own, lr=1e-2, eps=1e-15, norm_group=dist.group.WORLD)
        for _ in range(3):
            opt2.zero_grad(set_to_none=True)
            expert_loss(mine, rank).backward()
            opt2.step(max_norm=1.0)
        err = max(float((a - b).abs().max() / (b.abs().max() + 1e-12)) for a, b in zip(own, ref.submodules[rank].parameters()))
        untouched = all(torch.equal(a, b) for a, b in zip(mine.submodules[1 - rank].parameters(), fresh().submodules[1 - rank].parameters()))
        norm_err = abs(float(opt2.last_norm) - ref_norm) / ref_norm
        assert ref_norm > 1.0                                              # the clip was active
        # rank-strided mask generation + reduction
        g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "scene_boxes.npz")))
        images = [dict(H=int(g["HW"][i, 0]) // 4, W=int(g["HW"][i, 1]) // 4, intrinsics=g["intrinsics"][i] / 4.0, c2w=g["c2w"][i])
                  for i in (0, 50, 120, 200)]
        cen4, aabb = torch.from_numpy(g["centroids"]).to(dev), torch.from_numpy(g["aabb_global"]).to(dev)
        kw = dict(ray_samples=64, boundary_margin=float(g["margin"]), cluster_2d=True)
        full = voronoi_masks_and_boxes(images, cen4, aabb, **kw)
        part = voronoi_masks_and_boxes(images[rank::world], cen4, aabb, **kw)
        red = reduce_expert_aabbs(*part)
        same = all(torch.equal(a, b) for a, b in zip(red, full))
        fin = finalize_expert_boxes(*red, cen4, aabb)
        ret[rank] = (err, untouched, norm_err, same, bool(torch.isfinite(fin[0]).all() and (fin[1] >= fin[0]).all()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_per_expert_training_and_mask_generation_shard_across_gpus():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_modes, args=(2, _free_port(), ret), nprocs=2, join=True)
    for rank in range(2):
        err, untouched, norm_err, same, finite = ret[rank]
        assert err < 2e-5, (rank, err)              # the rank's expert after 3 globally-clipped Adam steps = the one-process run
        assert untouched                            # the other expert's replica is never written
        assert norm_err < 1e-5                      # sharded global norm = global norm
        assert same and finite                      # MIN / MAX / SUM reduction of the strided images = all images on one rank


def _worker_comm(rank, world, port, ret):
    """The C-ABI communicator (libacn_b200_comm.so, include/acn_b200_comm.h) driven with raw pointers as a C host would:
    the rendezvous id travels over a host channel (here a TCP store), then all-reduce, all-gather and the variable-size
    all-to-all of routed sample rows."""
    from torch.distributed import TCPStore
    from adaptive_city_nerf_b200 import comm
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    store = TCPStore("127.0.0.1", port, world, is_master=(rank == 0))
    if rank == 0:
        store.set("uid", comm.unique_id())
    c = comm.Communicator(dev, store.get("uid"), rank, world)
    try:
        x = torch.arange(1000, dtype=torch.float32, device=dev) * (rank + 1)
        c.allreduce_(x)
        ok = bool(torch.equal(x, torch.arange(1000, dtype=torch.float32, device=dev) * sum(r + 1 for r in range(world))))
        h = torch.full((7,), float(rank + 1), dtype=torch.float16, device=dev)
        c.allreduce_(h, comm.OP_MAX)
        ok &= bool((h == world).all())
        counts = torch.tensor([10 * rank + 3, rank + 1], dtype=torch.int32, device=dev)      # rows this rank sends to rank 0, 1
        allc = c.allgather(counts)                                                             # (world, world)
        ok &= allc.tolist() == [[10 * r + 3, r + 1] for r in range(world)]
        send_counts = counts.tolist()
        recv_counts = [int(allc[r, rank]) for r in range(world)]
        rows = torch.empty(sum(send_counts), 6, device=dev)
        off = 0
        for dst, n in enumerate(send_counts):           # row value encodes (source, destination, index)
            rows[off:off + n] = (100 * rank + 10 * dst) + torch.arange(n, device=dev, dtype=torch.float32)[:, None] / 64
            off += n
        got = c.alltoall_samples(rows, send_counts, recv_counts)
        off = 0
        for src, n in enumerate(recv_counts):
            want = (100 * src + 10 * rank) + torch.arange(n, device=dev, dtype=torch.float32)[:, None] / 64
            ok &= bool(torch.equal(got[off:off + n], want.expand(n, 6)))
            off += n
        torch.cuda.synchronize()
        ret[rank] = ok
    finally:
        c.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_c_abi_communicator_two_gpus():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_comm, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
