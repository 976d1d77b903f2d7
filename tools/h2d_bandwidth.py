"""Developer tool: aggregate host-to-device bandwidth of the box, one process per GPU (torch.distributed.run).

Every rank copies a pinned 256 MB buffer to its GPU with cudaMemcpyAsync (torch copy_, non_blocking) 20 times,
all ranks at once; prints per-rank and aggregate GB/s.  This is the ceiling of bench.py's `e2e` leg, which moves
215 MB of features per rank and step over PCIe (plus 41 MB of mask rows read in place).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_bandwidth.py
"""
import os

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = 256 << 20
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.fill_(rank + 1)
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for _ in range(3):
        dst.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    e0.record()
    for _ in range(iters):
        dst.copy_(host, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    gbs = torch.tensor([nbytes * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9], device=dev)
    all_gbs = [torch.zeros_like(gbs) for _ in range(world)]
    if world > 1:
        dist.all_gather(all_gbs, gbs)
    else:
        all_gbs = [gbs]
    if rank == 0:
        vals = [float(v.item()) for v in all_gbs]
        print(f"H2D from pinned memory, {world} rank(s) at once, {nbytes >> 20} MB x {iters}: per rank "
              + ", ".join(f"{v:.1f}" for v in vals) + f" GB/s; aggregate {sum(vals):.1f} GB/s", flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
