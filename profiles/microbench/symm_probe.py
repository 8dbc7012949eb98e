"""Probe (2+ GPUs, torchrun): does torch symmetric memory give peer pointers and an NVLS multicast pointer here?"""
import os
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
import torch.distributed._symmetric_memory as symm_mem
try:
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=torch.device("cuda", lr))
    h = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    print(rank, "buffer_ptrs", [hex(p) for p in h.buffer_ptrs], "multicast_ptr", hex(h.multicast_ptr) if h.multicast_ptr else None,
          "signal_pad_ptrs", len(h.signal_pad_ptrs), flush=True)
    t.fill_(float(rank + 1))
    h.barrier()
    peer = h.get_buffer((rank + 1) % world, (16,), torch.float32)
    print(rank, "peer read", peer[:2].tolist(), flush=True)
    h.barrier()
except Exception as e:
    import traceback; traceback.print_exc()
    print(rank, "FAILED", repr(e), flush=True)
print(rank, "can_access_peer", [torch.cuda.can_device_access_peer(lr, j) for j in range(world) if j != lr], flush=True)
dist.destroy_process_group()
