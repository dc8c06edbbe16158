"""N > 1 host logic on CPU: world_size-2 (and 3, uneven shards) gloo process groups exercise
the parameter sharding and the result gather that bench.py / multi-GPU callers use.  The
per-shard compute here is the CPU oracle (tests may use it); on GPUs it is the CUDA path."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, total, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mcmcdiag_b200 as m
    from oracle import mcmcdiag_oracle as o
    x = o.ar1(0.5, np.sqrt(0.75), 200, 4, total, rng=np.random.default_rng(3))   # same on every rank
    out = m.sharding.sharded_call(lambda xs: o.ess_rhat(xs), x)
    lo, hi = m.sharding.shard_range(total, rank, world)
    assert (hi - lo) in (total // world, total // world + 1)
    summ = m.sharding.sharded_call(lambda xs: o.summary(xs, fields=("mean", "ess_bulk", "rhat")), x)   # dict results
    single = m.sharding.sharded_call(lambda xs: o.rhat(xs), x)                                           # one array
    if rank == 0:
        S, R = o.ess_rhat(x)
        assert out.shape == (2, total)
        assert np.array_equal(out[0].numpy(), S) and np.array_equal(out[1].numpy(), R)   # independent of W
        want = o.summary(x, fields=("mean", "ess_bulk", "rhat"))
        assert list(summ) == list(want) and all(np.array_equal(summ[k].numpy(), want[k]) for k in want)
        assert np.array_equal(single.numpy(), o.rhat(x))
        open(os.path.join(tmp, f"ok{world}"), "w").write("ok")
    else:
        assert out is None and summ is None and single is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 10), (3, 11)])
def test_sharded_gather_gloo(tmp_path, world, total):
    port = 29511 + world
    mp.spawn(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / f"ok{world}").exists()


def test_shard_ranges_cover_and_balance():
    import mcmcdiag_b200 as m
    for total in (0, 1, 7, 10, 1_000_000):
        for world in (1, 2, 3, 4, 8):
            r = [m.sharding.shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = m.sharding.shard_sizes(total, world)
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        m.sharding.shard_range(10, 2, 2)
