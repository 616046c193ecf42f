"""Multi-GPU plumbing: one process per GPU, candidates sharded by contiguous index range, one
all-reduce(max) of the packed (score, index) key per BO step.

No data-path collective other than that 8-byte reduction exists: the posterior is replicated and every
rank applies the same deterministic float64 rank-one update.  `broadcast_posterior` is the optional
drift guard / late-joiner path the north star mentions (S, S_inv, b, m, R, m' packed, 3 d^2 + 3 d doubles).
"""
from __future__ import annotations

_SIGN = -(1 << 63)


def is_distributed() -> bool:
    try:
        import torch.distributed as td
        return td.is_available() and td.is_initialized() and td.get_world_size() > 1
    except Exception:
        return False


def rank_world():
    if not is_distributed():
        return 0, 1
    import torch.distributed as td
    return td.get_rank(), td.get_world_size()


def shard_range(total: int, rank: int, world: int):
    """Contiguous [start, stop) of `total` items for `rank`; the first total % world ranks get one extra."""
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def allreduce_best(key_tensor):
    """In-place max over ranks of the uint64 key stored in an int64 tensor.  The key's top bit is the top bit
    of the orderable score, so it is flipped to make signed max == unsigned max."""
    if not is_distributed():
        return key_tensor
    import torch.distributed as td
    key_tensor.bitwise_xor_(_SIGN)
    td.all_reduce(key_tensor, op=td.ReduceOp.MAX)
    key_tensor.bitwise_xor_(_SIGN)
    return key_tensor


def broadcast_posterior(engine, beta, src=0):
    """Replace every rank's posterior with rank `src`'s."""
    if not is_distributed():
        return
    import torch
    import torch.distributed as td
    buf = engine.blr_export()
    b = torch.tensor([float(beta or 0.0)], dtype=torch.float64, device=buf.device)
    td.broadcast(buf, src=src)
    td.broadcast(b, src=src)
    engine.blr_import(buf, float(b.item()))
