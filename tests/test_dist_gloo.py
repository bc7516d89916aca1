"""CPU, world_size=2, gloo: the N>1 path (batch sharding + one logits all-gather)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pytorchvideo_b200 import parallel as PAR


def _worker(rank, world, port, ret):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, lr, w = PAR.init_process_group("gloo")
    assert (r, w) == (rank, world)
    g = torch.Generator().manual_seed(7)
    clips = torch.rand((8, 3, 2, 4, 4), generator=g)
    slow_fast = [clips[:, :, ::2], clips]
    local = PAR.shard_batch(slow_fast, rank, world)
    assert local[0].shape[0] == 4 and local[1].shape[0] == 4
    # stand-in for the per-rank forward: any per-sample function (eval forward has no cross-sample coupling)
    logits = local[1].flatten(1)[:, :5] * 2.0 + local[0].flatten(1)[:, :5]
    full = PAR.gather_logits(logits, world)
    expect = clips.flatten(1)[:, :5] * 2.0 + clips[:, :, ::2].flatten(1)[:, :5]
    ok = bool(torch.allclose(full, expect))
    # unequal shards (7 clips over 2 ranks: 4 + 3): still ONE equal-sized all-gather, pad rows trimmed
    odd = clips[:7]
    lo, hi = PAR.shard_bounds(7, rank, world)
    assert hi - lo == (4 if rank == 0 else 3)
    mine = PAR.shard_batch(odd, rank, world)
    full7 = PAR.gather_logits(mine.flatten(1)[:, :5] * 3.0, world, total=7)
    ok = ok and full7.shape == (7, 5) and bool(torch.allclose(full7, odd.flatten(1)[:, :5] * 3.0))
    try:
        PAR.gather_logits(mine.flatten(1)[:1, :5], world, total=7)
        ok = False
    except RuntimeError:
        pass
    ret[rank] = ok
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29600 + (os.getpid() % 300)
    procs = [mp.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret[0] and ret[1]
