"""Probe: does torch.distributed._symmetric_memory give peer-mapped buffers on this box? (torchrun, 2+ GPUs)"""
import os, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem
t = symm_mem.empty((world, 1024), dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "rank", hdl.rank, "world", hdl.world_size, flush=True)
t.fill_(-1)
hdl.barrier()
# every rank writes its slot into every peer through the mapped pointer
for q in range(world):
    peer = hdl.get_buffer(q, (world, 1024), torch.float32)
    peer[rank].fill_(float(rank))
hdl.barrier()
print(rank, "got", t[:, 0].tolist(), flush=True)
assert t[:, 0].tolist() == [float(r) for r in range(world)]
dist.barrier(); dist.destroy_process_group()
