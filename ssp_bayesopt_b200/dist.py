"""Multi-GPU plumbing: one process per GPU, candidates sharded by contiguous index range, one
maximum over the ranks' packed (score, index) keys per BO step -- `reduce_best`: one kernel
over NVLink peer memory (csrc/sspbo_dist.cu), or an NCCL all-reduce(max) where peer mappings are not available.

No data-path exchange other than that 8-byte reduction exists: the posterior is replicated and every
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


def reduce_best(engine):
    """(score, index) of the best candidate over all ranks, on every rank, after the rank's scoring calls; synchronises once.
    One process per GPU of an NVLink node: the keys are exchanged by one kernel over peer memory (csrc/sspbo_dist.cu,
    `DeviceEngine.best_global`); if that cannot be set up on every rank (no peer access, SSPBO_DIST_EXCHANGE=0, more than 16
    ranks) the key goes through `allreduce_best` (NCCL) as before.  A single process: `engine.best()`."""
    if not is_distributed():
        return engine.best()
    if engine._exchange_on is None:
        engine.dist_connect()
    if engine._exchange_on:
        return engine.best_global()
    engine.settle()                                # local key final before it is combined across ranks
    allreduce_best(engine._best)
    return engine.best()


def allgather_best(key_tensor, shard_start: int):
    """Winner over ranks when the GLOBAL candidate index does not fit the key's 32 index bits (more than 2^32 candidates per
    step): every rank scores its shard with indices local to the shard (index0 = 0, shard < 2^32 rows) and the ranks exchange
    (local key, shard start) pairs -- 16 bytes per rank, SURVEY.md section 8e -- instead of max-reducing one key.
    Returns (score, global_index) with np.argmax's tie-break (lowest global index).  `key_tensor` is left untouched."""
    from . import _lib
    import torch
    if not is_distributed():
        s, i = _lib.key_unpack(int(key_tensor.item()))
        return s, shard_start + i
    import torch.distributed as td
    world = td.get_world_size()
    mine = torch.stack([key_tensor.reshape(()).to(torch.int64),
                        torch.tensor(int(shard_start), dtype=torch.int64, device=key_tensor.device)])
    parts = [torch.empty_like(mine) for _ in range(world)]
    td.all_gather(parts, mine)
    best = None
    for part in parts:
        key, start = part.cpu().tolist()
        key &= 0xFFFFFFFFFFFFFFFF
        if key == 0:                                          # a rank without candidates
            continue
        _, i = _lib.key_unpack(key)
        cand = (key >> 32, -(start + i))                      # orderable score bits (NaN counts as the maximum), then lowest global index
        if best is None or cand > best:
            best = (cand[0], cand[1], key)
    if best is None:
        return float("-inf"), -1
    return _lib.key_unpack(best[2])[0], -best[1]


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
