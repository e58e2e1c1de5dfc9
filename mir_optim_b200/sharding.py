"""How work is split over the GPUs of one box.  Pure host logic.

* batched problems (configs[1], [2], [4]): contiguous, even split of the batch; no data-path collective.
* one large problem (configs[3]): contiguous row blocks of J / y; x, J^T J, J^T r and the LM control state are
  replicated, and the only exchange is the all-reduce inside mir_optimize_least_squares_sharded_d.

`exchange_unique_id` shares the 128-byte NCCL id over an existing torch.distributed group (any backend), which
is the only thing the C ABI needs from the host runtime to build its communicator.
"""
from __future__ import annotations


def even_split(total: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of `total` items for `rank`: sizes differ by at most one, earlier ranks get the extra item."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("rank/world out of range")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def row_shard(m: int, rank: int, world: int, align: int = 32) -> tuple[int, int]:
    """Row block of the large problem for `rank`.  Blocks are multiples of `align` rows (the SYRK stage height) except
    the last, so no rank pads in the middle of the matrix."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("rank/world out of range")
    tiles = (m + align - 1) // align
    lo_t, hi_t = even_split(tiles, rank, world)
    return min(lo_t * align, m), min(hi_t * align, m)


def exchange_unique_id(engine, dist, rank: int, src: int = 0) -> bytes:
    """rank `src` creates the NCCL unique id through the library, everyone receives it over `dist` (torch.distributed)."""
    import torch
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == src:
        buf = torch.frombuffer(bytearray(engine.nccl_unique_id()), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        buf = buf.cuda()
    dist.broadcast(buf, src=src)
    return bytes(buf.cpu().numpy().tobytes())
