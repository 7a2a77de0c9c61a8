"""Dataset loaders with the reference's signatures (data/load_data.py:11-136).

    load_training_set(train_dir, dataset, seed, number_of_entire_training_instances,
                      feature_size, patch, patch_sizes) -> (X, Y)
    load_test_set(test_dir, test_idx, dataset, feature_size, patch, patch_sizes) -> (X, Y)

X is (rows, points, feature_size) Fourier features of the pixel-centre coordinates of a row
(image / patch), Y is (rows, points, channels).  One n-dimensional routine replaces the
reference's per-modality copies (data/image.py, audio.py, video.py, protein.py): a datum
(channels, *spatial) is cut into `patch_sizes` tiles in row-major tile order (the nested
x_idx / y_idx [/ t_idx] loops of the reference) and every tile gets the coordinate features
of `utils.to_grid_coordinates_and_features` + the Fourier block (data/image.py:24-27).

On-disk formats, as the reference reads them:
  cifar / kodak : a directory of image files, sorted by name (PNG/JPEG through PIL; `.npy`
                  arrays (C,H,W) or (H,W,C) in [0,1] are accepted too, for boxes without PIL);
                  portrait images are rotated to landscape (data/image.py:18-19);
                  cifar test batches are 500 consecutive files, kodak 1 (load_data.py:87-106)
  audio / video / protein : `<dir>/train_dataset.pkl`, `<dir>/test_dataset.pkl` = pickled list
                  of tensors (audio (1,L); video (T,3,H,W); protein (3,L)); protein test batches
                  are 1000 consecutive entries (load_data.py:126-134)
The dataset *preprocessors* of the reference (LibriSpeech / UCF-101 / PDB download and crop)
need network and are out of scope.
"""
from __future__ import annotations

import itertools
import os
import pickle

import numpy as np
import torch

from recombiner_b200.utils import fourier_features, to_grid_coordinates_and_features

_TEST_BATCH = {"cifar": 500, "kodak": 1, "protein": 1000}


def datum_pairs(datum: torch.Tensor, feature_size: int, patch: bool, patch_sizes):
    """(channels, *spatial) -> (X, Y): one row if not `patch`, else one row per tile."""
    def pair(t):
        coords, feats = to_grid_coordinates_and_features(t)
        return fourier_features(coords, feature_size), feats
    if not patch:
        x, y = pair(datum)
        return x[None], y[None]
    spatial = datum.shape[1:]
    counts = [s // p for s, p in zip(spatial, patch_sizes)]
    xs, ys = [], []
    for tile in itertools.product(*[range(c) for c in counts]):
        sl = (slice(None),) + tuple(slice(i * p, i * p + p) for i, p in zip(tile, patch_sizes))
        x, y = pair(datum[sl])
        xs.append(x)
        ys.append(y)
    return torch.stack(xs), torch.stack(ys)


def _pairs(data, feature_size, patch, patch_sizes):
    xy = [datum_pairs(d, feature_size, patch, patch_sizes) for d in data]
    return torch.cat([x for x, _ in xy], 0), torch.cat([y for _, y in xy], 0)


def read_image(path: str) -> torch.Tensor:
    """(C, H, W) float32 in [0,1], landscape."""
    if path.endswith(".npy"):
        a = torch.from_numpy(np.load(path).astype(np.float32))
        if a.ndim == 3 and a.shape[-1] in (1, 3) and a.shape[0] not in (1, 3):
            a = a.permute(2, 0, 1)
    else:
        from PIL import Image
        a = torch.from_numpy(np.asarray(Image.open(path), dtype=np.uint8).copy())
        a = (a[..., None] if a.ndim == 2 else a).permute(2, 0, 1).float() / 255
    if a.shape[1] > a.shape[2]:
        a = a.permute(0, 2, 1)
    return a.contiguous()


def _listing(directory: str):
    return [os.path.join(directory, n) for n in sorted(os.listdir(directory))]


def _pickled(directory: str, name: str):
    with open(os.path.join(directory, name), "rb") as f:
        return pickle.load(f)


def _as_datum(dataset: str, t: torch.Tensor) -> torch.Tensor:
    t = torch.as_tensor(t, dtype=torch.float32)
    return t.permute(1, 0, 2, 3) if dataset == "video" else t       # video is stored (T, C, H, W)


def _subsample(items, seed, n):
    """`n` entries without replacement under np.random.seed(seed) (load_data.py:27-31)."""
    n = min(len(items), n)
    idx = np.random.RandomState(seed).choice(len(items), n, False)
    np.random.seed(None)
    return [items[i] for i in idx]


def load_training_set(train_dir, dataset, seed, number_of_entire_training_instances, feature_size, patch, patch_sizes):
    if dataset in ("cifar", "kodak"):
        data = [read_image(p) for p in _subsample(_listing(train_dir), seed, number_of_entire_training_instances)]
    else:
        data = [_as_datum(dataset, t) for t in
                _subsample(_pickled(train_dir, "train_dataset.pkl"), seed, number_of_entire_training_instances)]
    return _pairs(data, feature_size, patch, patch_sizes)


def load_test_set(test_dir, test_idx, dataset, feature_size, patch, patch_sizes):
    if dataset in ("cifar", "kodak"):
        n = _TEST_BATCH[dataset]
        data = [read_image(p) for p in _listing(test_dir)[test_idx * n:test_idx * n + n]]
    else:
        items = _pickled(test_dir, "test_dataset.pkl")
        n = _TEST_BATCH.get(dataset)
        items = items[test_idx * n:test_idx * n + n] if n else [items[test_idx]]
        data = [_as_datum(dataset, t) for t in items]
    return _pairs(data, feature_size, patch, patch_sizes)
