"""Host-side helpers with the reference's `utils` names (utils.py:200-298).

Metrics and coordinate features are tiny numpy/torch host code (SURVEY §2 C8/C9:
parity yard-sticks and input producers, not kernels).
"""
from __future__ import annotations

import numpy as np
import torch


def count_layer_params(in_dim, out_dim):
    """Weights plus biases of one dense layer (utils.py:215-221)."""
    return out_dim * (in_dim + 1)


def count_net_params(in_dim, hidden_dims, out_dim):
    """Per-layer parameter counts of the INR and their running sum (utils.py:224-231)."""
    widths = [in_dim, *hidden_dims, out_dim]
    per_layer = [count_layer_params(a, b) for a, b in zip(widths[:-1], widths[1:])]
    return per_layer, np.cumsum(per_layer)


def _quantise8(img):
    return np.round(np.clip(img, 0, 1) * 255) / 255


def PSNR(original, compressed, round, max_value=1):
    """PSNR of one signal (or all patches of one signal), optionally after 8-bit
    rounding of the reconstruction (utils.py:234-242)."""
    rec = _quantise8(compressed) if round else compressed
    mse = np.mean((original - rec) ** 2)
    return (20 * np.log10(max_value / np.sqrt(mse))).item()


def batch_PSNR(original, compressed, round, max_value=1):
    """Per-datapoint PSNR over a batch (utils.py:245-254)."""
    n = original.shape[0]
    rec = _quantise8(compressed) if round else compressed
    mse = np.mean((original.reshape(n, -1) - rec.reshape(n, -1)) ** 2, axis=-1)
    return 20 * np.log10(max_value / np.sqrt(mse))


def batch_RMSD(original, compressed, scale_factor):
    """Per-structure RMSD of 3-D coordinates normalised by `scale_factor` (utils.py:257-260)."""
    n = original.shape[0]
    d2 = (original * scale_factor - compressed * scale_factor) ** 2
    return (d2.reshape(n, -1).mean(-1) * 3) ** 0.5


_METRICS = {
    "cifar": lambda a, b: batch_PSNR(a, b, round=True, max_value=1),
    "kodak": lambda a, b: PSNR(a, b, round=True, max_value=1),
    "video": lambda a, b: PSNR(a, b, round=True, max_value=1),
    "audio": lambda a, b: PSNR(a, b, round=False, max_value=1),
    "protein": lambda a, b: batch_RMSD(a, b, scale_factor=25),
}


def metric(original, compressed, dataset):
    """Distortion reported for a modality (utils.py:200-213)."""
    fn = _METRICS.get(dataset)
    return None if fn is None else fn(original, compressed)


def make_coord_grid(shape, range, device=None):
    """Pixel-centre coordinates of a grid, each axis mapped to `range` (utils.py:265-284)."""
    axes = []
    for i, s in enumerate(shape):
        lo, hi = range[i] if isinstance(range[0], (list, tuple)) else range
        axes.append(lo + (hi - lo) * ((0.5 + torch.arange(s, device=device)) / s))
    return torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1)


def to_grid_coordinates_and_features(datum):
    """(channels, *spatial) -> coordinates (points, d) in (-1,1) and features (points, channels)
    (utils.py:287-298)."""
    spatial = datum.shape[1:]
    coords = make_coord_grid(spatial, (-1, 1), device=datum.device).view(-1, len(spatial))
    return coords, datum.reshape(datum.shape[0], -1).T


def fourier_features(coords, feature_size):
    """[cos(pi x w), sin(pi x w)] with w = exp(linspace(0, ln 1024, feature_size/(2d)))
    -- the block every reference loader repeats (data/image.py:25-27)."""
    d = coords.shape[-1]
    w = torch.exp(torch.linspace(0, np.log(1024), feature_size // (2 * d), device=coords.device))
    arg = torch.matmul(coords.unsqueeze(-1), w.unsqueeze(0)).view(*coords.shape[:-1], -1)
    return torch.cat([torch.cos(np.pi * arg), torch.sin(np.pi * arg)], dim=-1)


# --------------------------------------------------------------------------- #
# functional glue of the reference, on the kernels (utils.py:4-198)
# --------------------------------------------------------------------------- #
_ENGINES = {}


def _engine_for(upsample_net, latent_dim, pixel_sizes, upsample_factors, patch, patch_nums, data_dim, device):
    """A FitEngine holding this upsampler's folded weights (cached per module object and weight version)."""
    from .engine import FitEngine
    scales = [getattr(upsample_net, f"up{i}").scale_factor for i in (1, 2, 3)]
    scales = [tuple(int(v) for v in f) if isinstance(f, (tuple, list)) else int(f) for f in scales]
    pads = [int(getattr(upsample_net, f"conv{i}").padding[0]) for i in (1, 2, 3)]
    versions = tuple(p._version for p in upsample_net.parameters())
    key = (id(upsample_net), tuple(pixel_sizes), tuple(upsample_factors), bool(patch), tuple(patch_nums) if patch else None, str(device))
    ent = _ENGINES.get(key)
    if ent is None or ent[1] != versions:
        in_dim = 34 if data_dim == 3 else 32
        eng = ent[0] if ent is not None else FitEngine([in_dim, 32, 32, 32, 3], data_dim, list(pixel_sizes), list(upsample_factors),
                                                       latent_dim, scales, pads, 30.0, device,
                                                       patch_nums=list(patch_nums) if patch else None)
        eng.set_mappings([torch.zeros(c, c) for c in eng.counts], upsample_net.state_dict())
        _ENGINES[key] = ent = (eng, versions)
    return ent[0]


def map_lpe_to_inr_inputs(upsample_net, latent_pe, latent_dim, pixel_sizes, upsample_factors, patch, patch_nums, data_dim):
    """Latent positional encodings -> per-pixel encodings, (data_num, sample_size, pixels, 16) (utils.py:4-120): the
    rows of a datum are stitched into one grid, upsampled as a whole by the folded polyphase kernels and cut back into
    patches.  latent_pe: (sample_size, data_num, L) or (sample_size, data_num, *grid, latent_dim).  Evaluation only: the
    result carries no autograd graph (the fit path differentiates through FitEngine, not through this function)."""
    if not latent_pe.is_cuda:
        from ._lib import KernelError
        raise KernelError("map_lpe_to_inr_inputs runs on the sm_100a kernels: pass CUDA tensors (no CPU fallback)")
    S, N = latent_pe.shape[:2]
    eng = _engine_for(upsample_net, latent_dim, pixel_sizes, upsample_factors, patch, patch_nums, data_dim, latent_pe.device)
    return eng.upsample_latents(latent_pe.detach().reshape(S, N, -1))


def map_hierarchical_model_to_int_weights(use_hierarchical_model, loc, scale, h_loc, h_scale, hh_loc, hh_scale, sample_size,
                                          hierarchical_patch_nums, patch_nums, data_dim):
    """h_w (data_num, sample_size, W) = mu + sigma * eps, plus the level-2 and level-3 terms of the patch hierarchy,
    each with its own per-patch noise (utils.py:122-198), drawn by rcb_fit_sample (Philox keyed from torch's global
    generator, so `torch.manual_seed` controls it).  scale arguments are standard deviations, as in the reference.
    Evaluation only (no autograd graph)."""
    import ctypes as C
    from . import _lib
    from ._lib import KernelError, SampleArgs, check, ptr, stream
    if not loc.is_cuda:
        raise KernelError("map_hierarchical_model_to_int_weights runs on the sm_100a kernels: pass CUDA tensors")
    lib = _lib.load()
    dev = loc.device
    N, W = loc.shape
    ld = (W + 3) // 4 * 4
    hw = torch.zeros(N * sample_size, ld, device=dev)
    seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
    levels = [(loc, scale, None)]
    if use_hierarchical_model:
        R = int(np.prod(patch_nums))
        l2 = hierarchical_patch_nums['level2']
        ng = [patch_nums[i] // l2[i] for i in range(data_dim)]
        n = np.arange(N)
        pc = np.unravel_index(n % R, patch_nums)
        grp = np.ravel_multi_index([pc[i] // l2[i] for i in range(data_dim)], ng)
        levels += [(h_loc, h_scale, (n // R) * int(np.prod(ng)) + grp), (hh_loc, hh_scale, n // R)]
    keep = []
    for li, (mu, sig, row_map) in enumerate(levels):
        mu = mu.detach().to(dev, torch.float32).contiguous()
        raw = torch.log(torch.expm1(6.0 * sig.detach().to(dev, torch.float32))).contiguous()      # inverse of softplus(.)/6
        a = SampleArgs()
        a.loc, a.log_scale, a.hw = ptr(mu), ptr(raw), ptr(hw)
        if row_map is not None:
            rm = torch.as_tensor(np.asarray(row_map).astype(np.int32), device=dev)
            keep.append(rm)
            a.row_map = ptr(rm)
        a.seed, a.row_offset = (42 << 32) | seed, 0
        a.rows, a.S, a.P, a.n_w, a.n_l, a.ld_hw = N, sample_size, W, W, 0, ld
        a.step, a.tensor_id, a.accumulate = 0, li, int(li > 0)
        keep += [mu, raw]
        check(lib.rcb_fit_sample(C.byref(a), stream()), "rcb_fit_sample")
    return hw[:, :W].reshape(N, sample_size, W).clone()
