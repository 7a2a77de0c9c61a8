"""The UNMODIFIED reference (oracle/_ref) driven through its own public API on the host cores.

TEST INFRASTRUCTURE / CPU BASELINE ONLY -- used by bench.py (`--impl reference`, `cpu_baseline`,
`e2e_short`) and by the trajectory tests; never by the product.  Nothing of recombiner_b200 is
on this path: the model is `test_model.TestBNNmodel(device='cpu')` of the copied reference, the
mappings are the reference's own `prior_model.LinearTransform` / `Upsample` modules, the block
grouping its own `get_grouping_by_kl`, and every step runs through `train` /
`optimize_posteriors` / `compress_posteriors` / `compress_group` exactly as
main_compression.py:148-162 calls them.
"""
from __future__ import annotations

import contextlib
import io
import time

import numpy as np
import torch

from . import build_ref


def mappings(ref, dims, data_dim, paddings, layer_scales, A, up_state):
    lt = ref.prior_model.LinearTransform(dims)
    with torch.no_grad():
        for p, a in zip(lt.A, A):
            p.copy_(a)
    up = ref.prior_model.Upsample(data_dim, paddings, layer_scales)
    up.load_state_dict({k: v.clone() for k, v in up_state.items()})
    return lt, up


def build_model(cfg, dataset, rows, A, up_state, p_loc, p_log_scale, grouping, init_log_scale=-4.0, initial_beta=1e-8,
                seed=42, quiet=True):
    """reference TestBNNmodel on the CPU for a single-level modality, as main_compression.py:93-146 builds it."""
    ref = build_ref.load()
    dims = [cfg["input_dim"]] + list(cfg["hidden_dims"]) + [cfg["output_dim"]]
    lt, up = mappings(ref, dims, cfg["data_dim"], cfg["paddings"], cfg["layerwise_scale_factors"], A, up_state)
    gi, gs, ge, g2p, p2g, G = grouping[:6]
    ctx = contextlib.redirect_stdout(io.StringIO()) if quiet else contextlib.nullcontext()
    with ctx:
        m = ref.test_model.TestBNNmodel(
            in_dim=cfg["input_dim"], hidden_dims=cfg["hidden_dims"], out_dim=cfg["output_dim"], number_of_datapoints=rows,
            upsample_factors=cfg["upsample_factors"], latent_dim=cfg["latent_dim"], data_dim=cfg["data_dim"],
            pixel_sizes=cfg["pixel_sizes"], patch=cfg["patch"], patch_nums=cfg["patch_nums"],
            hierarchical_patch_nums=cfg["hierarchical_patch_nums"], dataset=dataset, linear_transform=lt, upsample_net=up,
            p_loc=p_loc.clone(), p_log_scale=p_log_scale.clone(), init_log_scale=init_log_scale, param_to_group=p2g,
            group_to_param=g2p, n_groups=G, group_start_index=gs, group_end_index=ge, group_idx=gi,
            h_p_loc=None, h_p_log_scale=None, h_init_log_scale=None, h_param_to_group=None, h_group_to_param=None,
            h_n_groups=None, h_group_start_index=None, h_group_end_index=None, h_group_idx=None,
            hh_p_loc=None, hh_p_log_scale=None, hh_init_log_scale=None, hh_param_to_group=None, hh_group_to_param=None,
            hh_n_groups=None, hh_group_start_index=None, hh_group_end_index=None, hh_group_idx=None,
            w0=30., c=6., random_seed=seed, device="cpu", kl_upper_buffer=0., kl_lower_buffer=0.4, kl_adjust_gap=10,
            initial_beta=initial_beta, beta_step_size=0.05)
    return m


def time_fit(model, x, y, steps: int, warmup: int, lr: float = 2e-4, sample_size: int = 5):
    """Seconds per `TestBNNmodel.train` step (test_model.py:621-635), Adam as optimize_posteriors builds it."""
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    model.train(x=x, y=y, n_epochs=warmup, optimizer=opt, verbose=False, sample_size=sample_size)
    t0 = time.perf_counter()
    model.train(x=x, y=y, n_epochs=steps, optimizer=opt, verbose=False, sample_size=sample_size)
    return (time.perf_counter() - t0) / max(steps, 1)


def time_rec(model, pairs: int):
    """(seconds per candidate-table build, seconds per compress_group with a cached table)."""
    n = int(np.ceil(2 ** model.bit_per_group))
    blocks = list(range(min(pairs, model.n_groups)))
    t0 = time.perf_counter()
    for b in blocks:
        model.get_sample(b, n)                       # Sobol + norm.ppf, cached per block (test_model.py:459-471)
    t_table = (time.perf_counter() - t0) / len(blocks)
    t0 = time.perf_counter()
    for i, b in enumerate(blocks):
        model.sample_group(i % model.loc.shape[0], b, n)          # scoring only: leaves the model's state untouched
    t_pair = (time.perf_counter() - t0) / len(blocks)
    return t_table, t_pair


def short_schedule(model, x, y, n_fit: int, n_finetune: int, lr: float = 2e-4):
    """optimize_posteriors + compress_posteriors as main_compression.py:148-162 calls them, with a short
    schedule.  Returns (distortion array, wall seconds, coded KL bits per row)."""
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        model.optimize_posteriors(x, y, n_epochs=n_fit, lr=lr, verbose=0)
        kl_bits = model.update_annealing_factors(False).sum(1) / np.log(2.)
        d = model.compress_posteriors(x, y, n_epochs_finetune=n_finetune, h_n_epochs_finetune=None,
                                      hh_n_epochs_finetune=None, verbose=0, lr=lr, fine_tune_gap=1,
                                      compress_from_group_with_largest_kl=True)
    return np.asarray(d), time.perf_counter() - t0, kl_bits
