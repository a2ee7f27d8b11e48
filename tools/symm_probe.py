"""Probe torch symmetric memory on this box (2 ranks): attribute names, peer pointers, a peer read."""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
dev = torch.device("cuda", torch.cuda.current_device())
t = symm_mem.empty(4096, dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
if rank == 0:
    print([n for n in dir(hdl) if not n.startswith("_")])
    print("buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal_pad_ptrs", [hex(p) for p in hdl.signal_pad_ptrs],
          "signal_pad_size", hdl.signal_pad_size, "rank", hdl.rank, "world", hdl.world_size)
t.fill_(float(rank + 1))
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (4096,), torch.float32)
print(rank, "peer value", peer[:2].tolist())
hdl.barrier()
dist.destroy_process_group()
