"""world_size-2 gloo tests (CPU) of the N>1 host logic: row sharding, the all-reduced
f64 sufficient statistics of the EM prior update, and result gathering."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import recombiner_oracle as orc


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, rows, P, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from recombiner_b200 import parallel
    g = torch.Generator().manual_seed(3)
    loc = torch.randn(rows, P, generator=g) * 0.05
    log_scale = -4 + 0.5 * torch.randn(rows, P, generator=g)
    lo, hi = parallel.shard_rows(rows, world, rank)
    sl, ss = loc[lo:hi].double(), orc.std_transform(log_scale[lo:hi]).double()
    stats = torch.cat([sl.sum(0), (sl * sl).sum(0), (ss * ss).sum(0)])
    parallel.all_reduce_sum_(stats)
    mu, sc = parallel.prior_from_stats(stats, rows)
    ref_mu, ref_sc = orc.em_prior_update(loc, log_scale)
    np.testing.assert_allclose(mu.numpy(), ref_mu.numpy(), rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(sc.numpy(), ref_sc.numpy(), rtol=1e-5)
    gathered = parallel.gather_rows(loc[lo:hi])
    assert torch.equal(gathered, loc)
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_em_statistics_allreduce_world2(tmp_path):
    world = 2
    last = None
    for attempt in range(3):                      # a probed-free port can be taken before the workers bind it
        try:
            mp.spawn(_worker, args=(world, _free_port(), 11, 37, str(tmp_path)), nprocs=world, join=True)
            last = None
            break
        except Exception as e:                    # noqa: BLE001
            last = e
    assert last is None, last
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_shard_rows_partitions_exactly():
    from recombiner_b200.parallel import shard_rows
    for n, w, unit in ((1024, 8, 1), (1000, 8, 1), (7, 3, 1), (96 * 5, 4, 96), (60, 8, 60)):
        blocks = [shard_rows(n, w, r, unit) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        assert all((b - a) % unit == 0 for a, b in blocks)
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= unit
    with pytest.raises(ValueError):
        shard_rows(100, 4, 0, unit=96)
