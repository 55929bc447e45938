"""N > 1 host logic on CPU: world_size-2 gloo processes exercise env sharding and the end-of-rollout
statistics all-reduce (the only collective of the path)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    from marl_snake_b200 import shard_range
    for total in (1, 7, 64, 1048576, 1048577):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from marl_snake_b200 import allreduce_stats, shard_range
    lo, hi = shard_range(1001, rank, world)
    stats = torch.zeros(8, dtype=torch.float64)
    stats[0] = hi - lo            # pretend: episodes == envs owned
    stats[1] = float(sum(range(lo, hi)))
    allreduce_stats(stats)
    q.put((rank, stats.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_stats_allreduce_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = [q.get(timeout=100) for _ in ps]
    for p in ps:
        p.join(timeout=30)
        assert p.exitcode == 0
    for _, s in out:
        assert s[0] == 1001 and s[1] == float(sum(range(1001)))


def test_allreduce_is_noop_single_process():
    from marl_snake_b200 import allreduce_stats
    s = torch.arange(8, dtype=torch.float64)
    assert torch.equal(allreduce_stats(s.clone()), s)
