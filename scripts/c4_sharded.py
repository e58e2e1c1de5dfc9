"""configs[3] row-sharded over the ranks of one box.  Launch: torchrun --nproc-per-node N scripts/c4_sharded.py [--m M]
Every rank generates ITS rows, the library all-reduces [lower(J^T J), J^T r] and the trial ||r||^2 over NCCL.
Rank 0 prints one JSON line: timing (max over ranks, device-synchronised), result, cross-rank bit-identity of x."""
import argparse, json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mir_optim_b200 as mo
from mir_optim_b200 import workloads, sharding

ap = argparse.ArgumentParser()
ap.add_argument("--rows", dest="m", type=int, default=4_000_000)
ap.add_argument("--K", type=int, default=42)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--max-iterations", type=int, default=0)
ap.add_argument("--noise", type=float, default=1e-3)
ap.add_argument("--robust", action="store_true", help="robust termination: maxGoodResidual = 4 m noise^2 => fConverged (SURVEY 8c P3)")
ap.add_argument("--no-comm", action="store_true", help="each rank solves its own rows independently (diagnostic)")
ap.add_argument("--own-stream", action="store_true")
a = ap.parse_args()
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
eng = mo.engine
comm = None
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not a.no_comm:
        comm = eng.nccl_comm_init(world, sharding.exchange_unique_id(eng, dist, rank), rank)
lo, hi = sharding.row_shard(a.m, rank, world)
wl = workloads.c4_gaussmix(m=a.m, K=a.K, noise=a.noise, row_slice=(lo, hi))
t = torch.from_numpy(wl.t).cuda(); y = torch.from_numpy(wl.y).cuda()
s = eng.settings()
if a.max_iterations:
    s.maxIterations = a.max_iterations
if a.robust:
    s.maxGoodResidual = 4.0 * a.m * a.noise * a.noise
out = None
side = torch.cuda.Stream() if a.own_stream else None
for rep in range(a.reps):
    x = wl.x0[0].copy()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r, st = eng.optimize_sharded(s, wl.model, x, wl.l, wl.u, t, y, comm=comm, want_stats=True, stream=side.cuda_stream if side else None)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    xs = torch.from_numpy(x).cuda()
    same = True
    if world > 1 and not a.no_comm:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        g = [torch.empty_like(xs) for _ in range(world)]
        dist.all_gather(g, xs)
        same = all(torch.equal(g[0], gi) for gi in g)
    out = {"world": world, "m": a.m, "n": wl.n, "rows_this_rank": hi - lo, "solve_s": float(tt.item()), "status": r.status,
           "iterations": r.iterations, "fCalls": r.fCalls, "gCalls": r.gCalls, "residual": r.residual, "lambda": r.lambda_,
           "passes": st["passes"], "iterations_per_s": r.iterations / float(tt.item()), "x_bit_identical_across_ranks": same,
           "x": x.tolist()}
    if rank == 0 or a.no_comm:
        print(json.dumps({k: v for k, v in out.items() if k != "x"} if a.no_comm else out), flush=True)
if comm is not None:
    eng.nccl_comm_destroy(comm)
    dist.destroy_process_group()
