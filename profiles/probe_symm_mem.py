"""Probe: does torch symmetric memory give peer-mapped buffers on this box (2+ GPUs, NCCL group)?"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
g = dist.group.WORLD
try:
    t = symm_mem.empty(256, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, g.group_name)
    t.fill_(float(rank + 1))
    torch.cuda.synchronize(); dist.barrier()
    peers = [hdl.get_buffer(r, (256,), torch.float32) for r in range(hdl.world_size)]
    vals = [float(p[0]) for p in peers]
    print(f"rank {rank}: world {hdl.world_size} buffer_ptrs {[hex(p) for p in hdl.buffer_ptrs]} peer values {vals}", flush=True)
except Exception as e:
    print(f"rank {rank}: symmetric memory unavailable: {type(e).__name__}: {e}", flush=True)
dist.barrier()
dist.destroy_process_group()
