"""Receiver side (the reference stops at CSVs of indices): a packed
bitstream of the 16-bit block indices and a decoder that rebuilds the signal from
(prior checkpoint, bitstream, seed) alone -- no posterior, no target.

    python -m recombiner_b200.decode --dataset cifar --prior_path PRIOR.pkl --bitstream x.rcb --out recon.npy

Bitstream: magic 'RCB1', u32 rows, u32 n_levels, per level u32 level_rows + u32 n_blocks, then
the indices of every level (level 1 first) as little-endian u16, row-major.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import struct

import numpy as np
import torch

from .config import configs
from . import utils

MAGIC = b"RCB1"


def pack_bitstream(index_tables) -> bytes:
    """index_tables: list of (rows_l, G_l) arrays (level 1[, level 2, level 3]) of indices < 65536."""
    out = [MAGIC, struct.pack("<II", int(index_tables[0].shape[0]), len(index_tables))]
    for t in index_tables:
        out.append(struct.pack("<II", int(t.shape[0]), int(t.shape[1])))
    for t in index_tables:
        a = np.asarray(t)
        if a.min() < 0 or a.max() > 65535:
            raise ValueError("indices must fit 16 bits")
        out.append(a.astype("<u2").tobytes())
    return b"".join(out)


def unpack_bitstream(blob: bytes):
    if blob[:4] != MAGIC:
        raise ValueError("not an RCB1 bitstream")
    rows, n_levels = struct.unpack_from("<II", blob, 4)
    off = 12
    shapes = []
    for _ in range(n_levels):
        shapes.append(struct.unpack_from("<II", blob, off))
        off += 8
    tables = []
    for r, g in shapes:
        n = r * g
        tables.append(np.frombuffer(blob, dtype="<u2", count=n, offset=off).reshape(r, g).astype(np.int64))
        off += 2 * n
    return rows, tables


def bitstream_of(model) -> bytes:
    tabs = [model.compressed_idx_groupwise]
    if model.patch:
        tabs += [model.h_compressed_idx_groupwise, model.hh_compressed_idx_groupwise]
    return pack_bitstream(tabs)


def decode(dataset, prior_objects, blob, device="cuda", seed=42, x=None):
    """Rebuild the reconstruction (rows, pixels, out) from the bitstream.  Every block is
    regenerated from its index by rcb_rec_decode; the forward pass then runs with the fully
    coded (deterministic, sigma = 1e-15) posterior, exactly as the encoder's final predict."""
    from .main_compression import _level_kwargs
    from .test_model import TestBNNmodel
    config = configs[dataset]
    rows, tables = unpack_bitstream(blob)
    g1, p1, g2, p2, g3, p3, linear_transform, upsample_net = prior_objects
    kw = {}
    kw.update(_level_kwargs("", g1, p1, device))
    if config['patch']:
        kw.update(_level_kwargs("h_", g2, p2, device))
        kw.update(_level_kwargs("hh_", g3, p3, device))
    with contextlib.redirect_stdout(io.StringIO()):
        m = TestBNNmodel(in_dim=config['input_dim'], hidden_dims=config['hidden_dims'], out_dim=config['output_dim'],
                         number_of_datapoints=rows, upsample_factors=config['upsample_factors'],
                         latent_dim=config['latent_dim'], data_dim=config['data_dim'], pixel_sizes=config['pixel_sizes'],
                         patch=config['patch'], patch_nums=config['patch_nums'],
                         hierarchical_patch_nums=config['hierarchical_patch_nums'], dataset=dataset,
                         linear_transform=linear_transform.to(device), upsample_net=upsample_net.to(device),
                         random_seed=seed, device=device, initial_beta=p1[2],
                         layer_scales=config['layerwise_scale_factors'], paddings=config['paddings'], **kw)
    for li, (lv, tab) in enumerate(zip(m._levels, tables)):
        if tab.shape != (lv.rows, lv.G):
            raise ValueError(f"level {li}: bitstream has {tab.shape}, prior expects {(lv.rows, lv.G)}")
        lv.sample.copy_(m.decode_posteriors(tab, level=li))
        lv.mask.fill_(1.0)
        lv.coded.fill_(1)
        lv.idx.copy_(torch.as_tensor(tab.astype(np.int32), device=lv.device))
    if x is None:
        coords, _ = utils.to_grid_coordinates_and_features(torch.zeros(1, *config['pixel_sizes']))
        x = utils.fourier_features(coords, config['fourier_dim'])[None].expand(rows, -1, -1)
    with torch.no_grad():
        return m.predict(x.to(device), random_seed=0), m


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--dataset", choices=tuple(configs), required=True)
    ap.add_argument("--prior_path", required=True)
    ap.add_argument("--bitstream", required=True)
    ap.add_argument("--out", required=True, help=".npy file for the reconstruction (rows, pixels, channels)")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--device", default="cuda")
    args = ap.parse_args(argv)
    from .main_compression import load_prior
    y, _ = decode(args.dataset, load_prior(args.prior_path), open(args.bitstream, "rb").read(), args.device, args.seed)
    np.save(args.out, y.cpu().numpy())


if __name__ == "__main__":
    main()
