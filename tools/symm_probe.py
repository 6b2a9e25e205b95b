"""Feasibility probe for peer-memory kernels: torch symmetric memory rendezvous, peer pointers, remote store + barrier."""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
buf = symm.empty(1 << 20, dtype=torch.float32, device=dev)
hdl = symm.rendezvous(buf, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast", hdl.has_multicast_support, flush=True)
buf.fill_(float(rank))
hdl.barrier()
peer = (rank + 1) % world
remote = hdl.get_buffer(peer, (1 << 20,), torch.float32)
remote[: 1 << 19].fill_(100.0 + rank)          # remote stores into the peer's buffer
hdl.barrier()
torch.cuda.synchronize()
print(rank, "my buffer now:", float(buf[0]), float(buf[(1 << 19) + 5]), flush=True)
dist.destroy_process_group()
