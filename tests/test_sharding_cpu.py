"""Host-side sharding logic on CPU: split arithmetic, and the unique-id exchange over a world_size-2 gloo
group (the N > 1 plumbing without a GPU)."""
import os
import socket

import numpy as np
import pytest

from mir_optim_b200.sharding import even_split, row_shard, exchange_unique_id


@pytest.mark.parametrize("total,world", [(0, 1), (1, 4), (10, 3), (1 << 20, 8), (1000003, 7)])
def test_even_split_partitions(total, world):
    spans = [even_split(total, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == total
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("m,world", [(4_000_000, 8), (4_000_000, 3), (1000, 2), (31, 4), (0, 2), (1 << 22, 8)])
def test_row_shard_is_tile_aligned_partition(m, world):
    spans = [row_shard(m, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == m
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert all(lo % 32 == 0 for lo, _ in spans if lo < m)
    with pytest.raises(ValueError):
        row_shard(m, world, world)


def test_workload_row_slices_tile_the_full_problem():
    from mir_optim_b200 import workloads
    full = workloads.c4_gaussmix(m=200_000, K=3)
    parts = [workloads.c4_gaussmix(m=200_000, K=3, row_slice=row_shard(200_000, r, 3)) for r in range(3)]
    assert np.array_equal(np.concatenate([p.y for p in parts]), full.y) and np.array_equal(np.concatenate([p.t for p in parts]), full.t)
    assert all(np.array_equal(p.x0, full.x0) for p in parts)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class FakeEngine:                      # the real one asks NCCL; the exchange logic is what is under test
        def nccl_unique_id(self):
            return bytes(range(128))
    uid = exchange_unique_id(FakeEngine(), dist, rank)
    lo, hi = even_split(1 << 20, rank, world)
    import torch
    n = torch.tensor([hi - lo]); dist.all_reduce(n)
    q.put((rank, uid == bytes(range(128)), int(n.item())))
    dist.destroy_process_group()


def test_unique_id_exchange_and_batch_split_over_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue(); port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    got = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in ps]
    assert got == [(0, True, 1 << 20), (1, True, 1 << 20)]
