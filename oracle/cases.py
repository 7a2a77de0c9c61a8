"""Deterministic synthetic inputs shared by the golden generator and the tests.

TEST INFRASTRUCTURE ONLY (see oracle/recombiner_oracle.py header).  Everything is
derived from integer seeds with torch/numpy CPU generators, so the golden files
only have to hold *outputs*; `oracle/make_golden.py` (reference side, build
container) and `tests/` (oracle + CUDA side, any box with the same image) rebuild
identical inputs.  Workload recipe follows SURVEY.md §8(d) "Config 1".
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

from .recombiner_oracle import Shape, fourier_inputs, grouping_by_kl, layer_param_counts

# Reference modality shapes (config.py:28-137) plus three reduced patch shapes that
# exercise the stitching / hierarchy code of utils.py at fixture-friendly sizes.
SHAPES: Dict[str, dict] = {
    "cifar": dict(dims=[32, 32, 32, 32, 3], data_dim=2, pixel_sizes=[32, 32], upsample_factors=[16, 16], fourier_dim=16),
    "protein": dict(dims=[32, 32, 32, 32, 3], data_dim=1, pixel_sizes=[96], upsample_factors=[16], fourier_dim=16),
    "kodak": dict(dims=[32, 32, 32, 32, 3], data_dim=2, pixel_sizes=[64, 64], upsample_factors=[16, 16], fourier_dim=16,
                  patch=True, patch_nums=[8, 12], hier={"level2": [4, 4], "level3": [8, 12]}),
    "audio": dict(dims=[32, 32, 32, 32, 1], data_dim=1, pixel_sizes=[800], upsample_factors=[16], fourier_dim=16,
                  patch=True, patch_nums=[60], hier={"level2": [4], "level3": [60]}),
    "video": dict(dims=[34, 32, 32, 32, 3], data_dim=3, pixel_sizes=[24, 16, 16], upsample_factors=[24, 16, 16],
                  fourier_dim=18, patch=True, patch_nums=[1, 8, 8], hier={"level2": [1, 4, 4], "level3": [1, 8, 8]},
                  layer_scales=[(6, 4, 4), 2, 2]),
    # reduced patch shapes (not in config.py): same code paths, small tensors
    "patch2d": dict(dims=[32, 32, 32, 32, 3], data_dim=2, pixel_sizes=[16, 16], upsample_factors=[16, 16], fourier_dim=16,
                    patch=True, patch_nums=[2, 4], hier={"level2": [1, 2], "level3": [2, 4]}),
    "patch1d": dict(dims=[32, 32, 32, 32, 1], data_dim=1, pixel_sizes=[32], upsample_factors=[16], fourier_dim=16,
                    patch=True, patch_nums=[6], hier={"level2": [2], "level3": [6]}),
    "patch3d": dict(dims=[34, 32, 32, 32, 3], data_dim=3, pixel_sizes=[24, 16, 16], upsample_factors=[24, 16, 16],
                    fourier_dim=18, patch=True, patch_nums=[1, 2, 2], hier={"level2": [1, 1, 2], "level3": [1, 2, 2]},
                    layer_scales=[(6, 4, 4), 2, 2]),
}


def shape_of(name: str) -> Shape:
    kw = dict(SHAPES[name])
    kw.pop("fourier_dim")
    return Shape(**kw)


def fourier_dim_of(name: str) -> int:
    return SHAPES[name]["fourier_dim"]


def synthetic_bits(P: int, total_bits: float) -> np.ndarray:
    """Per-parameter KL bits: Gamma(2,1) draws rescaled to `total_bits`."""
    b = np.random.RandomState(0).gamma(2.0, 1.0, P)
    return b * (total_bits / b.sum())


def make_mappings(shape: Shape, seed: int = 42):
    """Random-init linear reparam matrices A_l (U(-1,1)/n, prior_model.py:19-21) and
    upsampler weights (torch default conv init), drawn from one CPU generator."""
    g = torch.Generator().manual_seed(seed)
    A = []
    for n in layer_param_counts(shape.dims):
        A.append((torch.rand(n, n, generator=g) * 2 - 1) / n)
    w_up = {}
    chans = [(128, 64, 5), (64, 64, 3), (64, 16, 3)]
    for i, (ci, co, k) in enumerate(chans, start=1):
        fan_in = ci * k ** shape.data_dim
        bound = 1.0 / math.sqrt(fan_in)
        w_up[f"conv{i}.weight"] = (torch.rand(co, ci, *([k] * shape.data_dim), generator=g) * 2 - 1) * bound
        w_up[f"conv{i}.bias"] = (torch.rand(co, generator=g) * 2 - 1) * bound
    return A, w_up


def make_level(rows: int, P: int, total_bits: float, seed: int, coded_frac: float = 0.0):
    """One level of posterior/prior state in group order with a synthetic grouping."""
    g = torch.Generator().manual_seed(seed)
    gi, gs, ge, g2p, p2g, G, _, _ = grouping_by_kl(synthetic_bits(P, total_bits))
    p_loc = torch.randn(P, generator=g) * 0.05
    p_log_scale = -2.0 + 0.3 * torch.randn(P, generator=g)
    loc = p_loc[None, :] + 0.02 * torch.randn(rows, P, generator=g)
    log_scale = -4.0 + 0.5 * torch.randn(rows, P, generator=g)
    coded = np.zeros((rows, G), dtype=bool)
    mask = torch.zeros(rows, P)
    sample = torch.zeros(rows, P)
    if coded_frac > 0:
        rs = np.random.RandomState(seed + 1)
        coded = rs.rand(rows, G) < coded_frac
        for r in range(rows):
            for b in np.nonzero(coded[r])[0]:
                mask[r, gs[b]:ge[b]] = 1.0
        sample = p_loc[None, :] + 0.03 * torch.randn(rows, P, generator=g)
        sample = sample * mask
    beta = 10.0 ** (-8 + 4 * torch.rand(rows, G, generator=g))
    beta = torch.where(torch.from_numpy(coded), torch.zeros_like(beta), beta)
    return dict(loc=loc, log_scale=log_scale, p_loc=p_loc, p_log_scale=p_log_scale,
                group_idx=gi, group_start=gs, group_end=ge, group_to_param=g2p, param_to_group=p2g,
                n_groups=G, coded=coded, mask=mask, sample=sample, beta=beta)


def make_fit_case(name: str, n_data: int, S: int, seed: int = 7, coded_frac: float = 0.25,
                  total_bits: float = 512.0) -> dict:
    """Inputs of one test-time fit step: targets, Fourier inputs, mappings, posterior
    state for every level and the noise tensors in the reference's draw order."""
    shape = shape_of(name)
    g = torch.Generator().manual_seed(seed)
    rows = n_data * (int(np.prod(shape.patch_nums)) if shape.patch else 1)
    W, L = shape.n_weights, shape.n_latent
    A, w_up = make_mappings(shape, seed=42)
    xf = fourier_inputs(shape.pixel_sizes, fourier_dim_of(name))
    x = xf[None].repeat(rows, 1, 1)
    y = torch.rand(rows, shape.n_pixels, shape.dims[-1], generator=g)
    case = dict(name=name, shape=shape, rows=rows, S=S, A=A, w_up=w_up, x=x, y=y)
    case["lvl1"] = make_level(rows, W + L, total_bits, seed + 10, coded_frac)
    eps = {"lpe": torch.randn(S, rows, L, generator=g), "w": torch.randn(rows, S, W, generator=g)}
    if shape.patch:
        r2 = rows // int(np.prod(shape.hier["level2"]))
        r3 = rows // int(np.prod(shape.hier["level3"]))
        case["lvl2"] = make_level(r2, W, total_bits / 4, seed + 20, coded_frac)
        case["lvl3"] = make_level(r3, W, total_bits / 4, seed + 30, coded_frac)
        eps["h"] = torch.randn(rows, S, W, generator=g)
        eps["hh"] = torch.randn(rows, S, W, generator=g)
    case["eps"] = eps
    return case


def make_prior_case(name: str, n_data: int, seed: int = 11) -> dict:
    """Inputs of one prior-training step (parameter order, S=1)."""
    shape = shape_of(name)
    g = torch.Generator().manual_seed(seed)
    rows = n_data * (int(np.prod(shape.patch_nums)) if shape.patch else 1)
    W = shape.n_weights
    A, w_up = make_mappings(shape, seed=42)
    xf = fourier_inputs(shape.pixel_sizes, fourier_dim_of(name))
    w_std = math.sqrt(6.0 / shape.dims[-2]) / 30.0
    case = dict(name=name, shape=shape, rows=rows, A=A, w_up=w_up,
                x=xf[None].repeat(rows, 1, 1),
                y=torch.rand(rows, shape.n_pixels, shape.dims[-1], generator=g),
                loc=torch.rand(rows, W, generator=g) * 2 * w_std - w_std,
                log_scale=-4.0 + 0.3 * torch.randn(rows, W, generator=g),
                lpe_loc=0.1 * torch.randn(rows, *shape.lpe_dims, shape.latent_dim, generator=g),
                lpe_log_scale=-4.0 + 0.3 * torch.randn(rows, *shape.lpe_dims, shape.latent_dim, generator=g))
    eps = {"lpe": torch.randn(rows, *shape.lpe_dims, shape.latent_dim, generator=g),
           "w": torch.randn(rows, 1, W, generator=g)}
    if shape.patch:
        r2 = rows // int(np.prod(shape.hier["level2"]))
        r3 = rows // int(np.prod(shape.hier["level3"]))
        for tag, r in (("h", r2), ("hh", r3)):
            case[f"{tag}_loc"] = torch.rand(r, W, generator=g) * 2 * w_std - w_std
            case[f"{tag}_log_scale"] = -4.0 + 0.3 * torch.randn(r, W, generator=g)
            eps[tag] = torch.randn(rows, 1, W, generator=g)
    case["eps"] = eps
    case["kl_beta"] = 1e-4
    sp = math.log1p(math.exp(-2.0)) / 6
    case["prior"] = dict(loc=torch.zeros(W), scale=torch.full((W,), sp),
                         lpe_loc=torch.zeros(*shape.lpe_dims, shape.latent_dim),
                         lpe_scale=torch.full((*shape.lpe_dims, shape.latent_dim), sp))
    return case


def make_rec_case(D: int, seed: int = 5) -> dict:
    """One (row, block) REC problem of block size D."""
    g = torch.Generator().manual_seed(seed * 1000 + D)
    p_loc = (torch.randn(D, generator=g) * 0.05).numpy()
    p_scale = (torch.rand(D, generator=g) * 0.03 + 0.005).numpy()
    q_loc = p_loc + (torch.randn(D, generator=g) * 0.02).numpy()
    q_scale = (p_scale * (0.2 + 0.6 * torch.rand(D, generator=g).numpy())).astype(np.float32)
    return dict(D=D, p_loc=p_loc.astype(np.float32), p_scale=p_scale.astype(np.float32),
                q_loc=q_loc.astype(np.float32), q_scale=q_scale)
