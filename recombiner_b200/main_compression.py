"""CLI: compress test datapoints with a trained prior.  Same flags, checkpoint stream
and CSV outputs as the reference driver (main_compression.py:12-178); the model runs on
the sm_100a kernels.  Extra, optional knobs (environment variables) shorten the schedule
for smoke runs without touching the reference defaults:
RECOMBINER_FIT_EPOCHS (30000), RECOMBINER_FINETUNE_EPOCHS (max(30000//G, 50))."""
from __future__ import annotations

import argparse
import os
import pickle

import numpy as np
import torch

from .config import configs
from .test_model import TestBNNmodel


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--seed', type=int, default=42)
    ap.add_argument('--test_dir', required=True)
    ap.add_argument('--test_idx', type=int, required=True)
    ap.add_argument("--dataset", choices=("cifar", "kodak", "video", "audio", "protein"))
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--prior_path", required=True, help='path of the learned prior, linear transform and upsampling net.')
    ap.add_argument("--save_dir", required=True, help='dir to save the compress files.')
    ap.add_argument("--save_bitstream", default=True)
    return ap.parse_args(argv)


def load_prior(path):
    """The 8-object pickle stream written by prior training (main_prior_training.py:284-335)."""
    with open(path, "rb") as f:
        return [pickle.load(f) for _ in range(8)]


def _level_kwargs(prefix, grouping, prior, device):
    """Reorder one level's prior into group order and go from sigma to raw scale,
    raw = ln(exp(6 sigma) - 1) (main_compression.py:49-66)."""
    group_idx, start, end, g2p, p2g, n_groups, _, _ = grouping
    p_loc, p_scale, _, avg_log_scale = prior
    if p_loc is None:
        return {}
    raw = torch.log(torch.exp(p_scale * 6) - 1)
    return {prefix + "p_loc": p_loc.clone()[p2g].to(device), prefix + "p_log_scale": raw.clone()[p2g].to(device),
            prefix + "init_log_scale": avg_log_scale[p2g].cpu().detach(),
            prefix + "param_to_group": p2g, prefix + "group_to_param": g2p, prefix + "n_groups": n_groups,
            prefix + "group_start_index": start, prefix + "group_end_index": end, prefix + "group_idx": group_idx}


def compress(x, y, dataset, prior_objects, device, seed=42, fit_epochs=None, finetune_epochs=None, verbose=1,
             row_offset=0, precision=None, code=True):
    """Library form of the driver body: returns (distortion, model).  `precision` selects the kernel family
    ('tf32' = tcgen05, the default; 'fp32' = SIMT parity path); `code=False` stops after optimize_posteriors
    (returns (None, model))."""
    config = configs[dataset]
    g1, p1, g2, p2, g3, p3, linear_transform, upsample_net = prior_objects
    kw = {}
    kw.update(_level_kwargs("", g1, p1, device))
    if config['patch']:
        kw.update(_level_kwargs("h_", g2, p2, device))
        kw.update(_level_kwargs("hh_", g3, p3, device))
    kl_beta = p1[2]
    model = TestBNNmodel(in_dim=config['input_dim'], hidden_dims=config['hidden_dims'], out_dim=config['output_dim'],
                         number_of_datapoints=x.shape[0], upsample_factors=config['upsample_factors'],
                         latent_dim=config['latent_dim'], data_dim=config['data_dim'], pixel_sizes=config['pixel_sizes'],
                         patch=config['patch'], patch_nums=config['patch_nums'],
                         hierarchical_patch_nums=config['hierarchical_patch_nums'], dataset=dataset,
                         linear_transform=linear_transform.to(device), upsample_net=upsample_net.to(device),
                         w0=30., c=6., random_seed=seed, device=device, kl_upper_buffer=0., kl_lower_buffer=0.4,
                         kl_adjust_gap=10, initial_beta=kl_beta, beta_step_size=0.05, row_offset=row_offset,
                         layer_scales=config['layerwise_scale_factors'], paddings=config['paddings'],
                         precision=precision, **kw).to(device)
    n_groups = kw["n_groups"]
    h_n, hh_n = kw.get("h_n_groups"), kw.get("hh_n_groups")
    short = finetune_epochs            # an explicit fine-tune length also shortens the level-2/3 rounds
    fit_epochs = int(os.environ.get("RECOMBINER_FIT_EPOCHS", 30000)) if fit_epochs is None else fit_epochs
    if finetune_epochs is None:
        finetune_epochs = int(os.environ.get("RECOMBINER_FINETUNE_EPOCHS", max(30000 // n_groups, 50)))
    x, y = x.to(device), y.to(device)
    model.optimize_posteriors(x, y, n_epochs=fit_epochs, lr=2e-4, verbose=verbose)
    if not code:
        return None, model
    distortion = model.compress_posteriors(
        x, y, n_epochs_finetune=finetune_epochs,
        h_n_epochs_finetune=None if h_n is None else (max(15000 // h_n, 20) if short is None else short),
        hh_n_epochs_finetune=None if hh_n is None else (max(15000 // hh_n, 20) if short is None else short),
        verbose=verbose, lr=2e-4, fine_tune_gap=1, compress_from_group_with_largest_kl=True)
    return distortion, model


def _write_csvs(args, config, distortion, tables):
    """The reference's output files (main_compression.py:163-178)."""
    if isinstance(distortion, float):
        distortion = np.array([[distortion]])
    np.savetxt(args.save_dir + "Distortion_test_id_%d" % args.test_idx + ".csv", distortion, delimiter=",")
    if int(args.save_bitstream):
        names = ("GroupIndex", "H_GroupIndex", "HH_GroupIndex")
        for name, t in zip(names, tables):
            np.savetxt(args.save_dir + name + "_test_id_%d" % args.test_idx + ".csv", t, delimiter=",")


def main(argv=None):
    """Single process: the reference driver.  Under `torchrun` (one process per GPU) the rows of the test batch are
    sharded by datapoint (by whole datum for the patch modalities) with no data-path collective; the per-row
    distortions and index tables are gathered at the end and rank 0 writes the same CSVs
    (main_compression.py:76-84 load, :163-178 write)."""
    import torch.distributed as dist
    from . import parallel
    args = parse_args(argv)
    config = configs[args.dataset]
    from data.load_data import load_test_set      # input producer outside the kernel path
    x, y = load_test_set(args.test_dir, args.test_idx, args.dataset, config['fourier_dim'], config['patch'],
                         config['pixel_sizes'])
    world = int(os.environ.get("WORLD_SIZE", "1"))
    device = args.device
    unit = int(np.prod(config['patch_nums'])) if config['patch'] else 1
    sharded = world > 1 and x.shape[0] // unit >= world
    if world > 1:
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        device = "cuda:%d" % local
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device(device))
        if not sharded and dist.get_rank() != 0:      # fewer data than GPUs (one kodak image): rank 0 works alone
            return
    lo, hi = parallel.shard_rows(x.shape[0], world, dist.get_rank(), unit) if sharded else (0, x.shape[0])
    distortion, model = compress(x[lo:hi], y[lo:hi], args.dataset, load_prior(args.prior_path), device, seed=args.seed,
                                 row_offset=lo, verbose=1 if (not sharded or dist.get_rank() == 0) else 0)
    tables = [model.compressed_idx_groupwise]
    if config['patch']:
        tables += [model.h_compressed_idx_groupwise, model.hh_compressed_idx_groupwise]
    if sharded:
        dev = torch.device(device)
        d = np.atleast_1d(np.asarray(distortion, dtype=np.float64)).reshape(-1)
        distortion = parallel.gather_rows(torch.from_numpy(d).to(dev)).cpu().numpy()
        tables = [parallel.gather_rows(torch.from_numpy(t).to(dev)).cpu().numpy() for t in tables]
        if not config['patch'] or distortion.size == 1:
            distortion = distortion if distortion.size > 1 else float(distortion[0])
        if dist.get_rank() != 0:
            return
    _write_csvs(args, config, distortion, tables)


if __name__ == '__main__':
    main()
