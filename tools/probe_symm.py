"""Probe: does torch symmetric memory rendezvous work on this box? (run under torchrun, 2+ GPUs)"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = symm.empty((1024,), dtype=torch.int32, device=dev)
t.fill_(rank + 1)
hdl = symm.rendezvous(t, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast", hex(hdl.multicast_ptr) if hdl.has_multicast_support else None,
      "signal", [hex(p) for p in hdl.signal_pad_ptrs][:2], flush=True)
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.int32)
print(rank, "peer value", int(peer[0].item()), flush=True)
hdl.barrier()
dist.destroy_process_group()
