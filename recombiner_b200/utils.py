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
