"""NCCL all_to_all_single / send-recv bandwidth between the ranks of one box (torchrun)."""
import os, sys, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
for mb in (16, 128, 512):
    n = mb * (1 << 20) // 4
    x = torch.randn(n, device=dev); y = torch.empty_like(x)
    for _ in range(3): dist.all_to_all_single(y, x)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): dist.all_to_all_single(y, x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    sent = mb * (world - 1) / world        # MiB leaving this rank
    if rank == 0: print(f"all_to_all_single {mb} MiB buffer: {ms:.3f} ms -> {sent / 1024 / (ms * 1e-3):.1f} GiB/s out per rank", flush=True)
    # all_reduce for comparison
    for _ in range(3): dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier()
    e0.record()
    for _ in range(10): dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    if rank == 0: print(f"all_reduce        {mb} MiB buffer: {ms:.3f} ms -> busbw {2 * (world - 1) / world * mb / 1024 / (ms * 1e-3):.1f} GiB/s", flush=True)
dist.destroy_process_group()
