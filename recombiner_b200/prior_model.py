"""Learned mappings, block grouping and the prior-training model.

Module-level names match the reference `prior_model` (prior_model.py:16-316) so
prior checkpoints -- which pickle `LinearTransform` and `Upsample` *module
objects* (main_prior_training.py:334-335) -- load unchanged.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn


class LinearTransform(nn.Module):
    """One dense square matrix A_l per INR layer, size out*(in+1); layer weights are
    `h_w[seg_l] @ A_l`.  Init U(-1,1)/n (prior_model.py:16-21)."""

    def __init__(self, net_dims):
        super().__init__()
        sizes = [net_dims[i] * (net_dims[i - 1] + 1) for i in range(1, len(net_dims))]
        self.A = nn.ParameterList([nn.Parameter((torch.rand(n, n) * 2 - 1) / n) for n in sizes])


class Upsample(nn.Module):
    """Parameter container of the latent upsampler: nearest-up x f1 -> conv(128->64, k5)
    -> LeakyReLU -> up x f2 -> conv(64->64, k3) -> LeakyReLU -> up x f3 -> conv(64->16, k3)
    (prior_model.py:23-59).  Attribute names (up1..3, conv1..3, act1..2) follow the
    reference so pickled checkpoints stay interchangeable.  The fit path never calls
    this module's forward: `FitEngine.set_mappings` folds the nearest-upsampling into
    the conv taps and runs the polyphase kernels instead."""

    def __init__(self, kernel_dim, paddings, layerwise_scale_factors):
        super().__init__()
        conv = {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}[kernel_dim]
        widths = [(128, 64, 5), (64, 64, 3), (64, 16, 3)]
        for i, ((cin, cout, k), pad, f) in enumerate(zip(widths, paddings, layerwise_scale_factors), start=1):
            setattr(self, f"up{i}", nn.Upsample(scale_factor=f))
            setattr(self, f"conv{i}", conv(cin, cout, k, padding=pad))
            if i < 3:
                setattr(self, f"act{i}", nn.LeakyReLU())
        self.kernel_dim = kernel_dim

    def forward(self, x):
        raise RuntimeError("Upsample.forward is not part of the B200 fit path: the upsampler runs inside "
                           "FitEngine as folded polyphase convolutions (rcb_upconv_fwd/bwd)")


# --------------------------------------------------------------------------- #
# block grouping (host side, prior_model.py:264-316)
# --------------------------------------------------------------------------- #
def group_parameters(parameters, weights, max_weight=16):
    """Greedy sequential binning: open a new block whenever adding the next
    parameter's KL would exceed `max_weight` bits (prior_model.py:301-316)."""
    groups, running = [[parameters[0]]], weights[0]
    for p, w in zip(parameters[1:], weights[1:]):
        if running + w > max_weight:
            groups.append([p])
            running = w
        else:
            groups[-1].append(p)
            running += w
    return groups


def get_grouping_by_kl(kls_bits):
    """Seed-0 shuffle of the parameters, then `group_parameters`; returns
    (group_idx, group_start_index, group_end_index, group2param, param2group, n_groups,
    group_kls, weights) as the reference does (prior_model.py:273-299)."""
    n = kls_bits.shape[0]
    order = np.random.RandomState(0).choice(n, n, False)     # == np.random.seed(0); np.random.choice
    np.random.seed(None)                                     # the reference also drops the global seed here
    groups = group_parameters(np.arange(n)[order], kls_bits[order])
    sizes = np.array([len(g) for g in groups])
    ends = np.cumsum(sizes)
    starts = ends - sizes
    param2group = np.concatenate([np.asarray(g) for g in groups])
    group2param = np.argsort(param2group)
    group_idx = np.repeat(np.arange(len(groups)), sizes).astype(int)
    group_kls = np.array([sum([kls_bits[i] for i in g]) for g in groups])
    return group_idx, starts, ends, group2param, param2group, len(groups), group_kls, kls_bits


def get_grouping(q_loc, q_scale, prior_loc, prior_scale):
    """Blocks from the training set's mean per-parameter KL in bits (prior_model.py:264-271)."""
    ratio = (q_scale / prior_scale) ** 2
    kl = 0.5 * (ratio + ((q_loc - prior_loc) / prior_scale) ** 2 - 1 - ratio.log())
    weights = (kl / np.log(2.)).mean(0).cpu().detach().numpy()
    return get_grouping_by_kl(weights)
