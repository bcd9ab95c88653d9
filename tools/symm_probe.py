"""Probe torch symmetric memory (peer pointers / NVLS multicast) on the GPU box. torchrun --nproc-per-node N tools/symm_probe.py"""
import os
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
import torch.distributed._symmetric_memory as symm_mem
print(rank, "symm_mem api:", [n for n in dir(symm_mem) if not n.startswith("_")][:40], flush=True)
try:
    t = symm_mem.empty((1 << 20,), dtype=torch.float64, device=torch.device("cuda", lr))
    h = symm_mem.rendezvous(t, dist.group.WORLD.group_name if hasattr(dist.group.WORLD, "group_name") else dist.group.WORLD)
    print(rank, "buffer_ptrs", [hex(p) for p in h.buffer_ptrs], "multicast_ptr", hex(h.multicast_ptr), "signal_pad_ptrs", len(h.signal_pad_ptrs), flush=True)
    t.fill_(rank + 1.0)
    h.barrier(channel=0)
    peer = h.get_buffer((rank + 1) % world, (8,), torch.float64)
    print(rank, "peer sees", peer[:2].tolist(), flush=True)
    h.barrier(channel=0)
except Exception as e:
    import traceback; traceback.print_exc()
    print(rank, "symm_mem FAILED:", repr(e), flush=True)
dist.destroy_process_group()
