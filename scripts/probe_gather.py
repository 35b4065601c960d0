"""Developer probe: where does the time of merge_shards go?  torchrun --nproc-per-node N scripts/probe_gather.py"""
import os, time, torch, torch.distributed as dist
rank, world, lrank = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(lrank); dev = torch.device("cuda", lrank)
dist.init_process_group("nccl", device_id=dev)
dist.all_reduce(torch.zeros(1, device=dev))
cap = (16384 // world) * 100
def T(name, fn, n=1):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter(); out = None
    for _ in range(n): out = fn()
    torch.cuda.synchronize()
    if rank == 0: print(f"{name:44s} {(time.perf_counter() - t0) * 1e3 / n:8.3f} ms", flush=True)
    return out
send = torch.zeros(cap, 32, dtype=torch.uint8, device=dev)
hdr = torch.zeros(512, dtype=torch.int64, device=dev)
counts = torch.zeros(world, dtype=torch.int64, device=dev)
T("all_gather counts (first)", lambda: dist.all_gather_into_tensor(counts, torch.tensor([cap], dtype=torch.int64, device=dev)))
T("all_gather counts (again)", lambda: dist.all_gather_into_tensor(counts, torch.tensor([cap], dtype=torch.int64, device=dev)))
T("all_reduce hdr int64[512] (first)", lambda: dist.all_reduce(hdr))
T("all_reduce hdr (again)", lambda: dist.all_reduce(hdr))
T("counts.tolist()", lambda: counts.tolist())
allrec = T("torch.empty(W*cap,32) first", lambda: torch.empty(world * cap, 32, dtype=torch.uint8, device=dev))
T("all_gather records (first)", lambda: dist.all_gather_into_tensor(allrec, send))
T("all_gather records (again, same buffers)", lambda: dist.all_gather_into_tensor(allrec, send), 3)
other = torch.empty(world * cap, 32, dtype=torch.uint8, device=dev)
T("all_gather records (new output buffer)", lambda: dist.all_gather_into_tensor(other, send))
T("torch.cat of the shards", lambda: torch.cat([allrec[r * cap:(r + 1) * cap] for r in range(world)]))
dist.destroy_process_group()
