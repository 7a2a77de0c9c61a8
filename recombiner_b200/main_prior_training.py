"""CLI + library form of prior training (main_prior_training.py:25-341): coordinate ascent
between (i) Adam on all posteriors and the shared mappings and (ii) the closed-form prior
update, with the global beta controller, and the reference's 8-object checkpoint stream.

Schedule constants default to the reference's (550 outer iterations x 200-then-100 epochs,
lr 2e-4, checkpoint every 10 iterations); `train_prior(...)` exposes them so short schedules
can be run.  With torch.distributed initialised each rank holds a contiguous shard of the
rows; ranks exchange only the shared-mapping gradients (per step) and the f64 sufficient
statistics + KL scalar (per outer iteration).
"""
from __future__ import annotations

import argparse
import pickle

import numpy as np
import torch
import torch.nn.functional as F

from . import parallel
from .config import configs
from .prior_model import LinearTransform, PriorBNNmodel, Upsample, em_prior_update, get_grouping, get_grouping_by_kl


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--seed', type=int, default=42, help='random seed')
    ap.add_argument('--train_dir', required=True, help='training dir')
    ap.add_argument('--train_size', type=int, default=10000000000)
    ap.add_argument("--dataset", choices=("cifar", "kodak", "video", "audio", "protein"))
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--max_bitrate", type=float, required=True)
    ap.add_argument("--saving_dir", default="./")
    return ap.parse_args(argv)


def budgets(dataset, config, max_bitrate):
    """KL budget window in bits per row (main_prior_training.py:75-83)."""
    pixels = np.prod(config['pixel_sizes'])
    scale = (3 / 48000) * 1000 if dataset == 'audio' else 1.0      # audio budgets are in kbps
    hi = max_bitrate * pixels * scale
    lo = max(config['lowest_bitrate'], max_bitrate - config['bitrate_range']) * pixels * scale
    assert lo <= hi
    return lo, hi


def step_beta(kl_beta, kl_bits, lo, hi):
    """x1.5 above the window, /1.5 below, clamped to [1e-20, 1] (main_prior_training.py:146-154)."""
    if kl_bits > hi:
        kl_beta *= 1.5
    if kl_bits < lo:
        kl_beta /= 1.5
    return min(max(kl_beta, 1e-20), 1)


def _grouping_over_ranks(q_loc, q_scale, p_loc, p_scale, n_total):
    """`get_grouping` (prior_model.py:264-271) with the per-parameter mean KL taken over the rows of ALL ranks:
    local sums, one all-reduce, then the same seed-0 greedy binning on every rank."""
    if parallel.world()[0] == 1:          # single process: the reference's own f32 arithmetic
        return get_grouping(q_loc, q_scale, p_loc, p_scale)
    ratio = (q_scale / p_scale) ** 2
    kl_bits = (0.5 * (ratio + ((q_loc - p_loc) / p_scale) ** 2 - 1 - ratio.log()) / np.log(2.)).sum(0).double()
    parallel.all_reduce_sum_(kl_bits)
    return get_grouping_by_kl((kl_bits / n_total).float().cpu().numpy())


def _mean_over_ranks(t, n_total):
    if parallel.world()[0] == 1:
        return t.detach().mean(0).cpu()
    s = t.detach().sum(0).double()
    parallel.all_reduce_sum_(s)
    return (s / n_total).float().cpu()


def checkpoint_objects(model, priors, kl_beta, linear_transform, upsample_net):
    """The 8 pickled objects of a prior checkpoint (main_prior_training.py:284-335).  Under torch.distributed the
    block grouping and the average training log-scale are statistics of the WHOLE training set (all-reduced sums
    over the ranks' shards), so every rank builds the same objects a single process would."""
    prior_loc, prior_scale, prior_lpe_loc, prior_lpe_scale = priors[:4]
    n1 = model.global_rows(0)
    with torch.no_grad():
        q_loc = torch.cat([model.loc.flatten(1), model.lpe_loc.flatten(1)], -1)
        q_scale = torch.cat([model.st(model.log_scale).flatten(1), model.st(model.lpe_log_scale).flatten(1)], -1)
        p_loc = torch.cat([prior_loc.flatten(), prior_lpe_loc.flatten()])
        p_scale = torch.cat([prior_scale.flatten(), prior_lpe_scale.flatten()])
        grouping = _grouping_over_ranks(q_loc, q_scale, p_loc, p_scale, n1)
        avg_ls = torch.cat([_mean_over_ranks(model.log_scale, n1), _mean_over_ranks(model.lpe_log_scale.flatten(1), n1)])
    none8 = (None,) * 8
    extra = [none8, (None, None, kl_beta, None), none8, (None, None, kl_beta, None)]
    if model.patch:
        extra = []
        for li, (q_l, q_ls) in enumerate(((model.h_loc, model.h_log_scale), (model.hh_loc, model.hh_log_scale))):
            pl, ps = priors[4 + 2 * li], priors[5 + 2 * li]
            n = model.global_rows(li + 1)
            with torch.no_grad():
                extra.append(_grouping_over_ranks(q_l, model.st(q_ls), pl, ps, n))
                extra.append((pl.cpu(), ps.cpu(), kl_beta, _mean_over_ranks(q_ls, n)))
    return [grouping, (p_loc.cpu(), p_scale.cpu(), kl_beta, avg_ls)] + extra + [linear_transform, upsample_net]


def save_checkpoint(path, objects):
    lt, up = objects[6], objects[7]
    dev = next(lt.parameters()).device
    with open(path, "wb") as f:
        for o in objects[:6]:
            pickle.dump(o, f)
        pickle.dump(lt.cpu(), f)
        pickle.dump(up.cpu(), f)
    lt.to(dev); up.to(dev)


def train_prior(X, Y, dataset, max_bitrate, device="cuda", seed=42, n_em_iter=550, first_epochs=200, epochs=100,
                lr=2e-4, checkpoint_every=10, on_checkpoint=None, verbose=True, row_offset=0, global_train_size=None):
    """Returns (checkpoint objects, list of ELBOs).  X, Y: this rank's rows."""
    config = configs[dataset]
    train_size = X.shape[0]
    model = PriorBNNmodel(in_dim=config['input_dim'], hidden_dims=config['hidden_dims'], out_dim=config['output_dim'],
                          train_size=train_size, data_dim=config['data_dim'], pixel_sizes=config['pixel_sizes'],
                          upsample_factors=config['upsample_factors'], latent_dim=config['latent_dim'],
                          patch=config['patch'], patch_nums=config['patch_nums'],
                          hierarchical_patch_nums=config['hierarchical_patch_nums'], random_seed=seed, device=device,
                          init_log_scale=-4, c=6., w0=30., layer_scales=config['layerwise_scale_factors'],
                          paddings=config['paddings'], row_offset=row_offset, global_train_size=global_train_size)
    linear_transform = LinearTransform(model.dims).to(device)     # same default init on every rank (same seed)
    upsample_net = Upsample(config['data_dim'], config['paddings'], config['layerwise_scale_factors']).to(device)
    world, rank = parallel.world()
    kl_beta = 1e-8
    lo, hi = budgets(dataset, config, max_bitrate)
    s0 = float(F.softplus(torch.tensor(-2.), beta=1, threshold=20) / 6)
    W, lpe_shape = model._W, model._lpe_shape
    priors = (torch.zeros(W, device=device), torch.full((W,), s0, device=device),
              torch.zeros(lpe_shape, device=device), torch.full(lpe_shape, s0, device=device))
    if config['patch']:
        priors += (torch.zeros(W, device=device), torch.full((W,), s0, device=device)) * 2
    X, Y = X.to(device), Y.to(device)
    elbos, objects, n_epoch = [], None, first_epochs
    for it in range(n_em_iter):
        p8 = tuple(priors) + (None,) * (8 - len(priors))
        _, kl_per_row, e = model.train(n_epoch, lr, X, Y, *p8, linear_transform, upsample_net, kl_beta,
                                       training_mappings=True, verbose=False)
        elbos += e
        n_epoch = epochs
        kl_bits = kl_per_row / np.log(2.)                 # average KL per row in bits (already all-reduced)
        kl_beta = step_beta(kl_beta, kl_bits, lo, hi)
        priors = em_prior_update(model)
        if it % checkpoint_every == 0 or it == n_em_iter - 1:
            if verbose and rank == 0:
                print("iter %d: Training KL %.4f bits/row; beta %.3g" % (it, kl_bits, kl_beta), flush=True)
            objects = checkpoint_objects(model, priors, kl_beta, linear_transform, upsample_net)
            if on_checkpoint is not None and rank == 0:
                on_checkpoint(objects, elbos)
    return objects, elbos, model


def main(argv=None):
    """Single process: the reference driver.  Under `torchrun` every rank loads the training set and keeps a
    contiguous shard of its rows (whole data for the patch modalities); rank 0 writes the checkpoints."""
    import os
    import torch.distributed as dist
    args = parse_args(argv)
    config = configs[args.dataset]
    from data.load_data import load_training_set     # input producer outside the kernel path
    n_inst = args.train_size // np.prod(config['patch_nums']) if config['patch'] else args.train_size
    X, Y = load_training_set(args.train_dir, args.dataset, args.seed, n_inst, config['fourier_dim'], config['patch'],
                             config['pixel_sizes'])
    train_size = X.shape[0]
    print("Prior is trained on %d patches/images." % train_size, flush=True)
    name = "_train_size_%d" % train_size + "_max_bitrate=%.3f.pkl" % args.max_bitrate
    device, lo, hi = args.device, 0, train_size
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        device = "cuda:%d" % local
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device(device))
        unit = int(np.prod(config['hierarchical_patch_nums']['level3'])) if config['patch'] else 1
        lo, hi = parallel.shard_rows(train_size, world, dist.get_rank(), unit)

    def write(objects, elbos):
        save_checkpoint(args.saving_dir + "PRIOR" + name, objects)
        with open(args.saving_dir + "LOSS" + name, "wb") as f:
            pickle.dump(elbos, f)

    train_prior(X[lo:hi], Y[lo:hi], args.dataset, args.max_bitrate, device=device, seed=args.seed, on_checkpoint=write,
                row_offset=lo, global_train_size=train_size,
                n_em_iter=int(os.environ.get("RECOMBINER_EM_ITERS", 550)),
                first_epochs=int(os.environ.get("RECOMBINER_FIRST_EPOCHS", 200)),
                epochs=int(os.environ.get("RECOMBINER_EPOCHS", 100)))


if __name__ == '__main__':
    main()
