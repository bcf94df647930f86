"""Platform probe: aggregate pinned-memory D2H / H2D bandwidth with 1, 2, 4, ... ranks copying at once (torch
copies only — none of the library's code). Launch under torch.distributed.run; rank 0 prints one JSON line."""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
nbytes = 256 << 20
dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
out = {"world": world, "mb_per_copy": nbytes >> 20}
active = 1
while active <= world:
    for name, (dst, src) in (("d2h", (host, dev)), ("h2d", (dev, host))):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if rank < active:
            for _ in range(20):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0 if rank < active else 0.0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        out[f"{name}_gbs_{active}_ranks"] = round(active * 20 * nbytes / float(dt.item()) / 1e9, 1)
    active *= 2
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
