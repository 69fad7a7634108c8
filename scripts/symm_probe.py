import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
t = symm.empty(1024, dtype=torch.float32, device=torch.device("cuda", rank))
h = symm.rendezvous(t, dist.group.WORLD)
t.fill_(float(rank + 1))
h.barrier()
peer = h.get_buffer((rank + 1) % world, (1024,), torch.float32)
print(rank, "peer value", float(peer[0]), "ptrs", [hex(p) for p in h.buffer_ptrs], "multicast", h.has_multicast_support(DeviceType := None) if False else None, flush=True)
h.barrier()
dist.destroy_process_group()
