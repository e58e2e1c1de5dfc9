"""All-reduce latency of the two message sizes the sharded LM path uses (8,384 doubles and 1 double), via torch.distributed."""
import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for n in (1, 8384, 1 << 20):
    buf = torch.ones(n, dtype=torch.float64, device="cuda")
    for _ in range(20): dist.all_reduce(buf)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): dist.all_reduce(buf)
    e1.record(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200): dist.all_reduce(buf)
    torch.cuda.synchronize(); host = (time.perf_counter() - t0) / 200
    if rank == 0: print(f"allreduce f64 x {n}: {e0.elapsed_time(e1) / 200 * 1e3:.1f} us/iter device, {host * 1e6:.1f} us/iter host-inclusive", flush=True)
print(rank, "can access peer", [torch.cuda.can_device_access_peer(local, j) for j in range(world) if j != local]) 
dist.destroy_process_group()
