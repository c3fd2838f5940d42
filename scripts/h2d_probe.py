"""Raw pinned host->device copy bandwidth on this box (the ceiling of every *_host entry point): one cudaMemcpyAsync per
buffer size, CUDA events, best of 5.  Run under torchrun to see the per-rank figure when N ranks copy at once."""
import json, os, sys, time
import torch
rank = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(rank)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
out = {}
for mb in (9.6, 38.5, 154.1):
    n = int(mb * 1e6)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    best = 1e9
    for rep in range(6):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        d.copy_(h, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        if rep:
            best = min(best, a.elapsed_time(b))
    out[f"{mb}MB"] = round(n / best / 1e6, 2)
print(json.dumps({"rank": rank, "world": world, "h2d_GBps": out}), flush=True)
