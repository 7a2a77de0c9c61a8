"""Relative entropy coding on device: candidate tables, Gumbel sequence, batched
encode / decode.  Host mirror of test_model.py:441-533,586-595.

The scrambled-Sobol state comes from `torch.quasirandom.SobolEngine` (the same
third-party engine the reference constructs at test_model.py:494); the points are
then produced on the GPU by random access from that state, so a receiver only
needs (D, seed) to rebuild the table of a block.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from ._lib import KernelError, RecArgs, check, ptr, stream

N_CAND_BITS = 16


def gumbel_sequence(seed: int, n: int) -> np.ndarray:
    """Sorted (decreasing) truncated-Gumbel noise shared by every block:
    with e_i ~ Exp(1) from numpy's legacy MT19937 stream, T_i = T_{i-1} + e_i (in the
    exp(-g) domain) and g_i = -ln T_i, each T re-derived from the rounded g as the
    reference does (test_model.py:446-456)."""
    expo = -np.log(np.random.RandomState(seed).rand(n))
    out = np.empty(n, dtype=np.float64)
    t_prev = 0.0
    for i in range(n):
        g = -np.log(expo[i] + t_prev)
        out[i] = g
        t_prev = np.exp(-g)
    return out


class CandidateTables:
    """Dimension-major f32 standard-normal candidate tables, one per distinct block size."""

    def __init__(self, seed: int, n_cand: int, device):
        self.seed, self.n, self.device = int(seed), int(n_cand), torch.device(device)
        self.lib = _lib.load()
        self._by_dim: Dict[int, torch.Tensor] = {}

    def table(self, D: int) -> torch.Tensor:
        t = self._by_dim.get(D)
        if t is None:
            eng = torch.quasirandom.SobolEngine(D, scramble=True, seed=self.seed)
            shift = eng.shift.to(self.device).contiguous()
            words = eng.sobolstate.to(self.device).contiguous()
            if words.shape != (D, 30) or words.dtype != torch.int64:
                raise KernelError("unexpected SobolEngine state layout")
            t = torch.empty(D, self.n, device=self.device, dtype=torch.float32)
            check(self.lib.rcb_rec_table(ptr(shift), ptr(words), ptr(t), D, self.n, stream()), "rcb_rec_table")
            self._by_dim[D] = t
        return t

    def pointer_array(self, sizes) -> torch.Tensor:
        """int64 device array: one table pointer per block."""
        return torch.tensor([self.table(int(d)).data_ptr() for d in sizes], dtype=torch.int64, device=self.device)


_WORKSPACES: Dict = {}


def _workspace(device, n_pairs: int) -> torch.Tensor:
    """Scratch of the staged scoring kernel, one per (device, stream): zero when allocated, left zero by every launch."""
    need = n_pairs * 772 + 128
    key = (str(device), stream())
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < need:      # a new (zeroed) buffer: the kernel's layout depends on the size only
        ws = _WORKSPACES[key] = torch.zeros(max(2 * need, 1 << 20), dtype=torch.uint8, device=device)
    return ws


def encode(lv, tables_ptr: torch.Tensor, gumbel: torch.Tensor, q_scale: torch.Tensor, p_scale: torch.Tensor,
           pair_row: torch.Tensor, pair_block: torch.Tensor, n_cand: int, max_D: int, *, apply: bool,
           want_logw: bool = False, staged: bool = True):
    """Code `len(pair_row)` (row, block) pairs.  apply=True commits the result to the
    level state (sample, mask, beta=0, coded, idx); otherwise returns (idx, z, log_w).
    staged=False forces the unstaged scoring kernel (same results bit for bit; kept for odd candidate counts)."""
    lib = _lib.load()
    n_pairs = int(pair_row.numel())
    dev = lv.device
    a = RecArgs()
    a.pair_row, a.pair_block = ptr(pair_row), ptr(pair_block)
    a.q_loc, a.q_scale, a.p_loc, a.p_scale = ptr(lv.loc.data), ptr(q_scale), ptr(lv.p_loc), ptr(p_scale)
    a.group_start, a.group_end, a.tables, a.gumbel = ptr(lv.group_start), ptr(lv.group_end), ptr(tables_ptr), ptr(gumbel)
    idx = z = logw = None
    if apply:
        a.idx_out, a.z_out = ptr(lv.idx), None
        a.sample, a.mask, a.beta, a.coded = ptr(lv.sample), ptr(lv.mask), ptr(lv.beta), ptr(lv.coded)
    else:
        idx = torch.empty(n_pairs, dtype=torch.int32, device=dev)
        z = torch.zeros(n_pairs, max_D, device=dev)
        a.idx_out, a.z_out = ptr(idx), ptr(z)
        a.sample = a.mask = a.beta = a.coded = None
    if want_logw:
        logw = torch.empty(n_pairs, n_cand, dtype=torch.float64, device=dev)
    a.logw_out = ptr(logw)
    a.n_pairs, a.P, a.G, a.n_cand, a.max_D, a.apply = n_pairs, lv.P, lv.G, n_cand, max_D, int(apply)
    if staged and os.environ.get("RECOMBINER_REC_STAGED", "1") != "0":
        ws = _workspace(dev, n_pairs)
        a.workspace, a.workspace_bytes = ptr(ws), ws.numel()
    check(lib.rcb_rec_encode(C.byref(a), stream()), "rcb_rec_encode")
    return idx, z, logw


def decode(lv, tables_ptr: torch.Tensor, p_scale: torch.Tensor, pair_row, pair_block, idx, n_cand: int,
           sample: torch.Tensor, mask: Optional[torch.Tensor]):
    """Receiver side: regenerate the coded values of the given (row, block, index) triples."""
    lib = _lib.load()
    check(lib.rcb_rec_decode(ptr(pair_row), ptr(pair_block), ptr(idx), ptr(lv.p_loc), ptr(p_scale),
                             ptr(lv.group_start), ptr(lv.group_end), ptr(tables_ptr), ptr(sample), ptr(mask),
                             int(pair_row.numel()), lv.P, n_cand, stream()), "rcb_rec_decode")
