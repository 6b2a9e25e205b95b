"""World-size-2 gloo tests (CPU) of the host-side multi-GPU logic in adaptive_city_nerf_b200/distributed.py:
count exchange, the variable-size sample round trip with autograd, expert regrouping, gradient all-reduce and the
sharded global-norm clip.  The per-sample arithmetic (routing, bucketing, fields, blending) is CUDA-only and is
covered by tests/test_gpu_multi.py; here the experts are stand-in torch modules."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_run, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


class _ToyExpert(torch.nn.Module):
    """Stand-in for MetaNGP: (M,6) -> (M,4), linear in its parameters so gradients are easy to predict."""

    def __init__(self, k):
        super().__init__()
        g = torch.Generator().manual_seed(100 + k)
        self.w = torch.nn.Parameter(torch.randn(6, 4, generator=g))

    def forward(self, rows):
        return rows[:, :6] @ self.w


def _toy_problem(rank, world, K=4, N=257):
    """Rows of one rank, bucketed the way ExpertShardedContainer does it on the device: grouped by destination rank,
    then by that rank's local expert (expert k lives on rank k % world)."""
    from adaptive_city_nerf_b200 import distributed as D
    g = torch.Generator().manual_seed(7 + rank)
    x = torch.randn(N, 6, generator=g)
    hard = torch.randint(0, K, (N,), generator=g)
    pos = torch.empty(K, dtype=torch.long)
    pos[torch.tensor(D.expert_send_order(K, world))] = torch.arange(K)
    order = torch.argsort(pos[hard], stable=True)
    counts = torch.bincount(hard, minlength=K)
    return x, hard, order, counts


def _exchange_worker(rank, world):
    from adaptive_city_nerf_b200 import distributed as D
    K = 4
    m = D.expert_owner_layout(K, world)
    experts = [_ToyExpert(k) for k in range(K)]                       # same seeds on every rank
    x, hard, order, counts = _toy_problem(rank, world, K)
    xd = x[order]
    owned = [e * world + rank for e in range(m)]
    local = [experts[k] for k in owned]
    all_counts = D.gather_counts(counts)
    y = D.routed_exchange(xd, all_counts, local)
    ref = torch.cat([experts[k](xd[hard[order] == k]) for k in D.expert_send_order(K, world)])
    err = float((y - ref).abs().max())
    # gradient of sum(y * c) w.r.t. every expert: owners receive the other rank's contribution through the backward exchange
    c = torch.arange(y.numel(), dtype=torch.float32).view_as(y) / y.numel()
    (y * c).sum().backward()
    grads = {k: experts[k].w.grad.clone() for k in owned}
    # reference: both ranks' data on one process
    full = [_ToyExpert(k) for k in range(K)]
    tot = 0.0
    for r in range(world):
        xr, hr, orr, _ = _toy_problem(r, world, K)
        yr = torch.cat([full[k](xr[orr][hr[orr] == k]) for k in D.expert_send_order(K, world)])
        cr = torch.arange(yr.numel(), dtype=torch.float32).view_as(yr) / yr.numel()
        tot = tot + (yr * cr).sum()
    tot.backward()
    gerr = max(float((grads[k] - full[k].w.grad).abs().max()) for k in grads)
    return err, gerr


def test_routed_exchange_roundtrip_and_gradients():
    for err, gerr in _spawn(_exchange_worker):
        assert err < 1e-6 and gerr < 1e-4, (err, gerr)


def _empty_worker(rank, world):
    """Rank 1 routes nothing and rank 0 routes everything to expert 0: the collectives must still line up."""
    from adaptive_city_nerf_b200 import distributed as D
    K = 2
    experts = [_ToyExpert(k) for k in range(K)]
    n = 33 if rank == 0 else 0
    xd = torch.randn(n, 6)
    counts = torch.tensor([n, 0])
    y = D.routed_exchange(xd, D.gather_counts(counts), [experts[rank]])
    (y.sum() * 1.0).backward()
    return tuple(y.shape), experts[rank].w.grad is not None


def test_routed_exchange_with_idle_ranks():
    (s0, g0), (s1, g1) = _spawn(_empty_worker)
    assert s0 == (33, 4) and s1 == (0, 4)
    assert g0 and not g1                       # only the owner of expert 0 saw rows


def _grads_worker(rank, world):
    from adaptive_city_nerf_b200 import distributed as D
    big = torch.nn.Parameter(torch.zeros(3000))
    small = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7))]
    big.grad = torch.full((3000,), float(rank + 1))
    small[0].grad = torch.full((5, 3), 10.0 * (rank + 1))
    small[1].grad = torch.full((7,), -1.0 * (rank + 1))
    D.allreduce_grads_([big] + small, average=True, flat_below=1000)
    ok = bool(torch.allclose(big.grad, torch.full((3000,), 1.5)) and torch.allclose(small[0].grad, torch.full((5, 3), 15.0))
              and torch.allclose(small[1].grad, torch.full((7,), -1.5)))
    # sharded clip: rank r owns a gradient of norm 3 (r=0) / 4 (r=1); a shared one of norm 12 is replicated
    own = torch.nn.Parameter(torch.zeros(1)); own.grad = torch.tensor([3.0 if rank == 0 else 4.0])
    sh = torch.nn.Parameter(torch.zeros(1)); sh.grad = torch.tensor([12.0])
    total = D.sharded_clip_grad_norm_([own], [sh], max_norm=1.0)
    mins = torch.tensor([[float(rank), 0.0, -float(rank)]]); maxs = mins + 1.0; cnt = torch.tensor([rank + 1])
    D.reduce_expert_aabbs(mins, maxs, cnt)
    return ok, float(total), float(own.grad), float(sh.grad), mins.tolist(), maxs.tolist(), int(cnt)


def test_allreduce_grads_clip_and_aabbs():
    for rank, (ok, total, og, sg, mins, maxs, cnt) in enumerate(_spawn(_grads_worker)):
        assert ok
        assert abs(total - 13.0) < 1e-5                                   # sqrt(3^2 + 4^2 + 12^2)
        assert abs(og - (3.0 if rank == 0 else 4.0) / 13.0) < 1e-5 and abs(sg - 12.0 / 13.0) < 1e-5
        assert mins == [[0.0, 0.0, -1.0]] and maxs == [[2.0, 1.0, 1.0]] and cnt == 3


def test_exchange_segments_layout():
    from adaptive_city_nerf_b200 import distributed as D
    assert D.expert_send_order(4, 2) == [0, 2, 1, 3]                 # rank 0 owns {0, 2}, rank 1 owns {1, 3}
    all_counts = torch.tensor([[5, 1, 0, 2], [3, 4, 7, 0]])          # rows (rank) x experts
    send, recv, segs = D.exchange_segments(all_counts, rank=0)
    assert send == [5 + 0, 1 + 2]                                     # to rank 0: experts 0, 2; to rank 1: experts 1, 3
    assert recv == [5 + 0, 3 + 7]                                     # from rank 0 / rank 1, my experts 0 and 2
    assert segs == [(0, 0, 5), (1, 5, 0), (0, 5, 3), (1, 8, 7)]       # (local expert, start, length) in buffer order
    send1, recv1, _ = D.exchange_segments(all_counts, rank=1)
    assert send1 == [3 + 7, 4 + 0] and recv1 == [1 + 2, 4 + 0]
    with pytest.raises(ValueError):
        D.expert_owner_layout(6, 4)
    assert D.expert_owner_layout(8, 4) == 2
