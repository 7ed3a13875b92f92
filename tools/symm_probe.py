"""Does torch's symmetric memory (peer-mapped buffers, NVLS multicast) work on this box?  2+ ranks under torchrun."""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = symm.empty(1 << 20, dtype=torch.float32, device=dev)
t.fill_(rank + 1.0)
h = symm.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "world", h.world_size, "multicast", h.has_multicast_support, hex(h.multicast_ptr) if h.has_multicast_support else None,
      "buffers", [hex(p) for p in h.buffer_ptrs], "signal", [hex(p) for p in h.signal_pad_ptrs], "pad bytes", h.signal_pad_size, flush=True)
h.barrier()
peer = h.get_buffer((rank + 1) % h.world_size, (1 << 20,), torch.float32)
print(rank, "peer value", float(peer[0]), flush=True)
h.barrier()
dist.barrier()
dist.destroy_process_group()
