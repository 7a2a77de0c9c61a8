"""Two-rank checks: datapoint-sharded compression reproduces the single-process result bit for bit (Philox counters
are keyed by the *global* row), and datum-sharded prior training with its all-reduces matches single-process training.
With two GPUs the ranks use one GPU each over NCCL; on a single-GPU box both ranks share cuda:0 and exchange through
gloo (NCCL refuses two ranks on one device) -- the sharding logic, the row offsets and every kernel are the same, only
the transport of the two all-reduces differs, so the test runs wherever the GPU suite runs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _case(rows):
    from oracle import cases
    case = cases.make_fit_case("cifar", rows, 5, coded_frac=0.0, total_bits=128.0)
    return case


def _compress(case, lo, hi, device):
    from tests.helpers import product_test_model
    sub = dict(case)
    sub["rows"] = hi - lo
    sub["lvl1"] = {k: (v[lo:hi] if (torch.is_tensor(v) or isinstance(v, np.ndarray)) and getattr(v, "shape", (0,))[:1] == (case["rows"],) else v)
                   for k, v in case["lvl1"].items()}
    m = product_test_model(sub, "cifar", device=device, precision="tf32")
    m.row_offset = lo
    x, y = case["x"][lo:hi].to(device), case["y"][lo:hi].to(device)
    m.optimize_posteriors(x, y, n_epochs=12, lr=2e-4, verbose=0)
    m.compress_posteriors(x, y, n_epochs_finetune=2, verbose=False, lr=2e-4)
    return m.compressed_idx_groupwise, m._lv.sample.cpu()


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    one_each = torch.cuda.device_count() >= world
    dev = rank if one_each else 0
    torch.cuda.set_device(dev)
    if one_each:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from recombiner_b200 import parallel
    case = _case(8)
    lo, hi = parallel.shard_rows(8, world, rank)
    idx, sample = _compress(case, lo, hi, f"cuda:{dev}")
    np.save(os.path.join(out_dir, f"idx{rank}.npy"), idx)
    torch.save(sample, os.path.join(out_dir, f"sample{rank}.pt"))
    # prior training on a shard, gradients of the shared mappings all-reduced over NCCL
    from recombiner_b200 import main_prior_training
    x, y = case["x"], case["y"]
    torch.manual_seed(0)
    objs, elbos, model = main_prior_training.train_prior(x[lo:hi], y[lo:hi], "cifar", 0.5, device=f"cuda:{dev}", n_em_iter=2,
                                                         first_epochs=3, epochs=2, checkpoint_every=1, verbose=False,
                                                         row_offset=lo, global_train_size=8)
    if rank == 0:
        torch.save(dict(A0=objs[6].A[0].detach().cpu(), p_loc=objs[1][0], p_scale=objs[1][1], elbo=elbos),
                   os.path.join(out_dir, "prior_dist.pt"))
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    case = _case(8)
    idx_ref, sample_ref = _compress(case, 0, 8, "cuda:0")
    idx = np.concatenate([np.load(tmp_path / f"idx{r}.npy") for r in range(world)])
    sample = torch.cat([torch.load(tmp_path / f"sample{r}.pt") for r in range(world)])
    np.testing.assert_array_equal(idx, idx_ref)                 # same indices as the unsharded run
    assert torch.equal(sample, sample_ref)
    from recombiner_b200 import main_prior_training
    torch.manual_seed(0)
    objs, elbos, _ = main_prior_training.train_prior(case["x"], case["y"], "cifar", 0.5, device="cuda:0", n_em_iter=2,
                                                     first_epochs=3, epochs=2, checkpoint_every=1, verbose=False)
    d = torch.load(tmp_path / "prior_dist.pt")
    # 5 Adam steps move every entry by ~5 * lr = 1e-3; the all-reduce changes the summation order
    # of the shared gradients, so compare at a tenth of that update
    np.testing.assert_allclose(d["A0"].numpy(), objs[6].A[0].detach().cpu().numpy(), rtol=0, atol=1e-4)
    np.testing.assert_allclose(d["p_loc"].numpy(), objs[1][0].numpy(), rtol=0, atol=1e-4)
    np.testing.assert_allclose(d["p_scale"].numpy(), objs[1][1].numpy(), rtol=1e-2)
    np.testing.assert_allclose(d["elbo"], elbos, rtol=1e-3)
