"""CPU: host logic of the batch sharding with world_size 2 over gloo -- partition, ordered gather of scalars."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from syngular_b200 import parallel


def test_shard_bounds_partition():
    for total in (0, 1, 7, 8, 8192, 8191):
        for world in (1, 2, 3, 4, 8):
            blocks = [parallel.shard_bounds(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            for (a, b), (c, d) in zip(blocks, blocks[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == total
    with pytest.raises(ValueError):
        parallel.shard_bounds(10, 2, 2)


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    r, w = parallel.init_from_env(backend="gloo")
    lo, hi = parallel.shard_bounds(total, r, w)
    local = torch.arange(lo, hi, dtype=torch.float64) * 1.5 + 0.25        # stands for the per-chain overlaps of this rank
    full = parallel.gather_scalars(local, total)
    ok = torch.equal(full, torch.arange(total, dtype=torch.float64) * 1.5 + 0.25)
    mx = parallel.max_over_ranks(10.0 + r, torch.device("cpu"))
    q.put((r, bool(ok), mx))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_gather_scalars_world2_gloo(total):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert [r[2] for r in res] == [11.0, 11.0]
