"""Data-parallel inference plumbing: one process per GPU, clips sharded along the batch dim,
ONE all-gather of the [B_local, num_classes] logits per step.  The forward
itself has no cross-sample coupling (eval-mode BatchNorm), so there is no other collective."""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_process_group(backend=None):
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_bounds(n_items, rank, world):
    """Contiguous balanced shard [lo, hi) of n_items for `rank` (first n%world ranks get one more)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(x, rank, world):
    """Rank-local slice of a clip batch (a tensor, or the [slow, fast] list of SlowFast)."""
    if isinstance(x, (list, tuple)):
        return [shard_batch(t, rank, world) for t in x]
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def gather_logits(local_logits, world=None, total=None):
    """ONE all_gather of the rank-local [B_local, K] logits -> [B_total, K] on every rank.

    ``total`` = global number of clips when the shards are unequal (``shard_bounds`` gives the first
    B %% world ranks one clip more): every rank then pads its rows to ceil(total / world) so that the
    collective is still a single equal-sized all_gather_into_tensor, and the pad rows are trimmed with the
    same ``shard_bounds``.  With ``total=None`` all ranks must hold the same number of rows."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return local_logits
    rows = local_logits.shape[0]
    if total is not None:
        rank = dist.get_rank()
        lo, hi = shard_bounds(total, rank, world)
        if hi - lo != rows:
            raise RuntimeError("rank %d holds %d rows but shard_bounds(%d, %d, %d) gives %d" % (rank, rows, total, rank, world, hi - lo))
        rows = -(-total // world)
    tail = tuple(local_logits.shape[1:])
    src = local_logits.contiguous()
    if rows != src.shape[0]:
        pad = torch.zeros((rows,) + tail, dtype=src.dtype, device=src.device)
        pad[: src.shape[0]] = src
        src = pad
    out = torch.empty((world * rows,) + tail, dtype=src.dtype, device=src.device)
    dist.all_gather_into_tensor(out, src)
    if total is None or total == world * rows:
        return out
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(total, r, world)
        parts.append(out[r * rows: r * rows + (hi - lo)])
    return torch.cat(parts, 0)
