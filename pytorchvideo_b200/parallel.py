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


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_cpus(local_rank, sysfs="/sys"):
    """(numa node, CPUs of that node) the GPU ``local_rank`` hangs off, from sysfs; (None, None) when unknown."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(os.path.join(sysfs, "bus/pci/devices", bus, "numa_node")).read())
        if node < 0:
            return None, None
        return node, _parse_cpulist(open(os.path.join(sysfs, "devices/system/node/node%d/cpulist" % node)).read())
    except Exception:
        return None, None


def bind_to_gpu_numa(local_rank):
    """One process per GPU: run this rank on the CPUs of its GPU's NUMA node BEFORE it allocates pinned host buffers, so
    the staging memory is first-touched on the socket whose PCIe root the GPU hangs off.  (Measured on an 8-GPU box with
    UNBOUND ranks, profiles/r02_scale.md: resident rate 7.94x of one GPU, end to end with f16 host clips 7.93x, with fp32
    host clips - 8 x 193 MB per 3 ms step - only 4.8x: host-memory / socket-interconnect bound.)  Returns the node, or
    None when sysfs does not say (no-op)."""
    node, cpus = gpu_numa_cpus(local_rank)
    if node is None or not cpus or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except OSError:
        return None


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
