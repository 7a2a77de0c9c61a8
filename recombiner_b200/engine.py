"""Host-side composition of the sm_100a kernels for the fit path.

`FitEngine` owns the device-side constants derived from the learned mappings
(padded / transposed reparameterisation matrices, upsample-folded conv weights)
and the per-(rows, S) workspaces, and sequences the C-ABI calls of one forward /
backward / update.  All arithmetic happens in librecombiner_b200.so; torch is
used for device memory and streams only.

Reference call sites replaced: test_model.py:283-355 (predict), :357-377
(calculate_kl), :621-635 (train step); prior_model.py:129-179 for the S=1 case.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import (KernelError, MlpArgs, ReduceArgs, SampleArgs, UpconvGeom, UpdateArgs, check, ptr, stream)


def _round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def layer_param_counts(dims: Sequence[int]) -> List[int]:
    return [dims[i + 1] * (dims[i] + 1) for i in range(len(dims) - 1)]


def _i32(a, device) -> torch.Tensor:
    return torch.as_tensor(np.ascontiguousarray(np.asarray(a)).astype(np.int32), device=device)


@dataclass
class Noise:
    """Either explicit eps tensors (parity tests) or a Philox key (production)."""
    eps_w: Optional[torch.Tensor] = None      # (rows, S, W)   level-1 weights
    eps_l: Optional[torch.Tensor] = None      # (S, rows, L)   latent grid
    seed: int = 0
    step: int = 0
    row_offset: int = 0
    eps_h: Optional[torch.Tensor] = None      # (rows, S, W)   level 2 (patch modalities)
    eps_hh: Optional[torch.Tensor] = None     # (rows, S, W)   level 3

    def eps_for(self, level: int):
        return (self.eps_w, self.eps_h, self.eps_hh)[level]


class LevelState:
    """One level of posterior state in group order plus its block structure, on device.
    Mirrors test_model.py:131-141,169,211-224 (level 1) / :144-166 (levels 2, 3)."""

    def __init__(self, loc, log_scale, p_loc, p_log_scale, group_idx, group_start, group_end,
                 group_to_param, param_to_group, beta, device):
        self.device = device
        self.loc = loc
        self.log_scale = log_scale
        self.p_loc = p_loc.detach().to(device=device, dtype=torch.float32).contiguous()
        self.p_log_scale = p_log_scale.detach().to(device=device, dtype=torch.float32).contiguous()
        self.p_scale_direct = False
        self.beta_scalar = 0.0
        self.rows, self.P = loc.shape
        self.adam = None
        # hierarchy (patch modalities): which level this is, per-column row permutation and
        # the expansion of level rows over patch rows
        self.level = 0
        self.perm = self.perm_inv = self.row_map = self.row_children = None
        self.n_children = 1
        if group_start is None:          # parameter order, no blocks (prior training)
            self.G = 0
            self.group_idx = self.group_start = self.group_end = self.g2p = self.p2g = None
            self.mask = self.sample = self.beta = self.coded = self.idx = self.group_kl = None
            return
        self.G = int(len(group_start))
        self.group_idx = _i32(group_idx, device)
        self.group_start = _i32(group_start, device)
        self.group_end = _i32(group_end, device)
        self.g2p = _i32(group_to_param, device)
        self.p2g = _i32(param_to_group, device)
        self.group_start_host = np.asarray(group_start).astype(np.int64)
        self.group_end_host = np.asarray(group_end).astype(np.int64)
        self.mask = torch.zeros(self.rows, self.P, device=device)
        self.sample = torch.zeros(self.rows, self.P, device=device)
        if torch.is_tensor(beta):
            beta = beta.detach().to(device=device, dtype=torch.float32)
            self.beta = (beta.expand(self.rows, self.G) if beta.ndim else beta.repeat(self.rows, self.G)).contiguous()
        else:
            self.beta = torch.full((self.rows, self.G), float(beta), device=device)
        self.coded = torch.zeros(self.rows, self.G, dtype=torch.uint8, device=device)
        self.idx = torch.zeros(self.rows, self.G, dtype=torch.int32, device=device)
        self.group_kl = torch.zeros(self.rows, self.G, dtype=torch.float64, device=device)

    def set_permutation(self, perm_g2p: np.ndarray):
        """perm_g2p[r, c] = stored row read by parameter-order row r in column c
        (test_model.py:185-194); the inverse serves the gradient scatter."""
        self.perm = _i32(perm_g2p, self.device)
        self.perm_inv = _i32(np.argsort(perm_g2p, axis=0), self.device)

    def set_expansion(self, row_map: np.ndarray):
        """row_map[n] = row of this level feeding patch row n (utils.py:151-189)."""
        row_map = np.asarray(row_map).astype(np.int64)
        self.row_map = _i32(row_map, self.device)
        order = np.argsort(row_map, kind="stable")
        counts = np.bincount(row_map, minlength=self.rows)
        assert counts.min() == counts.max(), "every level row must feed the same number of patches"
        self.n_children = int(counts[0])
        self.row_children = _i32(order.reshape(self.rows, self.n_children), self.device)

    def reset_adam(self):
        if getattr(self, "adam", None) is not None:      # in place: captured fit-step graphs keep these pointers
            for k in ("m1_loc", "v_loc", "m1_ls", "v_ls"):
                self.adam[k].zero_()
            self.adam["t"] = 0
            return
        z = lambda: torch.zeros(self.rows, self.P, device=self.device)
        self.adam = dict(m1_loc=z(), v_loc=z(), m1_ls=z(), v_ls=z(), t=0)


class SectionTimer:
    """CUDA-event brackets around named kernel launches on the launching stream
    (bench.py uses this for the per-kernel roofline; disabled by default)."""

    def __init__(self):
        self.events = {}

    def begin(self, name):
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        self.events.setdefault(name, []).append(ev)
        ev[0].record()
        return ev

    @staticmethod
    def end(ev):
        ev[1].record()

    def summary(self):
        torch.cuda.synchronize()
        return {k: (len(v), sum(a.elapsed_time(b) for a, b in v) / len(v)) for k, v in self.events.items()}


class _Section:
    def __init__(self, timer, name):
        self.timer, self.name = timer, name

    def __enter__(self):
        self.ev = self.timer.begin(self.name) if self.timer is not None else None

    def __exit__(self, *exc):
        if self.ev is not None:
            SectionTimer.end(self.ev)
        return False


class FitEngine:
    """Kernel sequencing for one modality shape (non-patch modalities: cifar, protein)."""
    timer: Optional[SectionTimer] = None
    map_generation = 0                            # bumped by set_mappings(): invalidates captured steps
    step_state: Optional[torch.Tensor] = None     # set while a fit step is being captured into a CUDA graph

    def section(self, name):
        return _Section(self.timer, name)

    @property
    def half_hw(self):
        """fp16 weight samples: tensor-core path, single-level posterior (nothing accumulates into hw), and no
        prior training (its dA_l = hw^T d_wt reads hw in fp32)."""
        return bool(self.tc and self.half_acts and self.patch_nums is None and all(o % 8 == 0 for o in self.offsets)
                    and len(getattr(self, "AT_h", ())) == len(self.counts))      # fp16 copies staged by set_mappings

    def __init__(self, dims, data_dim, pixel_sizes, upsample_factors, latent_dim, layer_scales, paddings,
                 w0, device, precision=None, patch_nums=None, force_poly=False):
        import os
        self.precision = precision or os.environ.get("RECOMBINER_PRECISION", "tf32")
        if self.precision not in ("fp32", "tf32"):
            raise KernelError(f"unknown precision {self.precision!r}: 'fp32' (SIMT parity path) or 'tf32' (tcgen05)")
        self.tc = self.precision == "tf32"
        self.tc_conv = self.tc and os.environ.get("RECOMBINER_TC_CONV", "1") != "0"
        self.tc_mlp = self.tc and os.environ.get("RECOMBINER_TC_MLP", "1") != "0"
        # the reparameterisation GEMMs (hw <-> wt) and the upsampler chain (lpe <-> pe) are independent
        # between the sampling kernel and the MLP: run them on two streams
        self.overlap = os.environ.get("RECOMBINER_OVERLAP", "1") != "0"
        # the four per-layer reparameterisation GEMMs as one launch: "1" both directions, "fwd" forward only, "0" off
        self.batch_gemm = os.environ.get("RECOMBINER_BATCH_GEMM", "fwd")
        self.half_dwt = os.environ.get("RECOMBINER_HALF_DWT", "1") != "0"
        self.half_pe = os.environ.get("RECOMBINER_HALF_PE", "1") != "0"
        # Fourier inputs regenerated from the pixel index inside the tensor-core MLP when the caller's x is the canonical one
        self.gen_x = os.environ.get("RECOMBINER_GEN_X", "1") != "0"
        self.x_generated = False
        # conv2's activations are only ever read as MMA operands (conv3) and for their signs (LeakyReLU mask):
        # where conv3 has the fp16-operand kernel they are stored as fp16 -- the 10 mantissa bits a TF32 MMA
        # reads anyway.  Prior training turns this off (its weight gradients read them in fp32).
        self.half_acts = self.tc_conv and os.environ.get("RECOMBINER_HALF_ACTS", "1") != "0"
        self.f2_half = False
        self.b2w = False
        self._b2w_enabled = os.environ.get("RECOMBINER_BWD_F2W", "1") != "0"
        self._half_dpe = os.environ.get("RECOMBINER_HALF_DPE", "1") != "0"
        self._fast_update = os.environ.get("RECOMBINER_FAST_UPDATE", "1") != "0"
        self._half_da1 = os.environ.get("RECOMBINER_HALF_DA1", "1") != "0"
        self.M1_h = None
        self._half_staged = False
        self._side = None
        if not torch.cuda.is_available():
            raise KernelError("recombiner_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device)
        _lib.use_device(self.device)          # 'cuda:1' as the reference accepts it: launches follow the model's device
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.dims = list(dims)
        self.counts = layer_param_counts(self.dims)
        self.offsets = [0] + list(np.cumsum(self.counts))[:-1]
        self.offsets = [int(o) for o in self.offsets]
        self.W = int(sum(self.counts))
        self.ldw = _round_up(self.W, 4)
        self.ldh = _round_up(self.W, 8)               # row stride of the fp16 weight samples
        if len(self.dims) != 5 or self.dims[1:4] != [32, 32, 32]:
            raise KernelError(f"unsupported INR {self.dims}: kernels are built for 3 hidden layers of 32")
        self.data_dim = data_dim
        if data_dim not in (1, 2, 3):
            raise KernelError("signals must be 1-D, 2-D or 3-D")
        self.pixel_sizes = list(pixel_sizes)
        self.pix = int(np.prod(pixel_sizes))
        self.grid = [pixel_sizes[i] // upsample_factors[i] for i in range(data_dim)]       # per-row latent grid
        self.latent_dim = int(latent_dim)
        self.L = int(np.prod(self.grid)) * self.latent_dim
        # patch modalities: the rows of one datum are stitched into one grid before the upsampler
        self.patch_nums = list(patch_nums) if patch_nums is not None else None
        pn = self.patch_nums or [1] * data_dim
        self.R = int(np.prod(pn))                                    # rows per datum
        self.full_grid = [pn[i] * self.grid[i] for i in range(data_dim)]
        self.full_pixels = [pn[i] * pixel_sizes[i] for i in range(data_dim)]
        self.sp = int(np.prod(self.grid))
        self.sp_total = int(np.prod(self.full_grid))
        self.pix_total = int(np.prod(self.full_pixels))
        self.lpe_slot = None
        self._patch_origin = None
        if self.patch_nums is not None:
            slot = np.empty((self.R, self.sp), dtype=np.int64)
            origin = np.empty(self.R, dtype=np.int64)
            for r in range(self.R):
                pc = np.unravel_index(r, pn)
                for sp_i in range(self.sp):
                    gc = np.unravel_index(sp_i, self.grid)
                    slot[r, sp_i] = np.ravel_multi_index([pc[i] * self.grid[i] + gc[i] for i in range(data_dim)], self.full_grid)
                origin[r] = np.ravel_multi_index([pc[i] * pixel_sizes[i] for i in range(data_dim)], self.full_pixels)
            self.lpe_slot = _i32(slot, torch.device(device))
            self._patch_origin = origin
        self.n_f = self.dims[0] - 16
        self.out = self.dims[-1]
        self.w0 = float(w0)
        if list(paddings) != [2, 1, 1]:
            raise KernelError("upsampler kernels assume 'same' convolutions (paddings [2,1,1])")
        # stage geometry: (d, h, w) grids of the three nearest-up + conv stages (leading axes
        # of lower-dimensional signals are singleton with factor 1 and kernel extent 1)
        pad = 3 - data_dim
        grid3 = [1] * pad + list(self.full_grid)
        ks = [5, 3, 3]
        chans = [(latent_dim, 64), (64, 64), (64, 16)]
        self.geoms = []
        for i in range(3):
            f = layer_scales[i]
            f = [int(v) for v in f] if isinstance(f, (tuple, list)) else [int(f)] * data_dim
            f3 = [1] * pad + f
            k3 = [1] * pad + [ks[i]] * data_dim
            self.geoms.append(UpconvGeom(*grid3, *f3, *k3, chans[i][0], chans[i][1]))
            grid3 = [grid3[j] * f3[j] for j in range(3)]
        if int(np.prod(grid3)) != self.pix_total or grid3[pad:] != self.full_pixels:
            raise KernelError("upsample factors do not reach the pixel grid")
        g0 = self.geoms[0]
        self.dense1 = data_dim < 3 and g0.h * g0.w <= 16 and not force_poly
        fp = [1] * pad + list(self.full_pixels)
        pp = [1] * pad + list(self.pixel_sizes)
        self.ph, self.pw = pp[1], pp[2]
        self.pitch_y, self.pitch_z = fp[2], fp[1] * fp[2]
        self._ws: Dict = {}
        self._x_cache = None
        self._xt_bufs = {}
        self.A = None

    # ---------------------------------------------------------------- mappings --
    def set_mappings(self, A_list, up_state):
        """Stage the learned mappings on device: zero-padded A_l and A_l^T, and the upsampler with its
        nearest-upsampling folded into the conv taps.  Buffers are allocated on the first call and refreshed in place
        afterwards (prior training re-stages every step; captured steps keep the addresses)."""
        conv_w = [up_state[f"conv{i}.weight"].detach().to(device=self.device, dtype=torch.float32).contiguous() for i in (1, 2, 3)]
        conv_b = [up_state[f"conv{i}.bias"].detach().to(device=self.device, dtype=torch.float32).contiguous() for i in (1, 2, 3)]
        first = self.A is None
        if first:
            self.A = [torch.zeros(c, _round_up(c, 4), device=self.device) for c in self.counts]
            self.conv_w = [torch.empty_like(w) for w in conv_w]
            self.conv_b = [torch.empty_like(b) for b in conv_b]
        for l, (a, c) in enumerate(zip(A_list, self.counts)):
            self.A[l][:, :c].copy_(a.detach())
        for dst, src in zip(self.conv_w + self.conv_b, conv_w + conv_b):
            dst.copy_(src)
        self.restage_mappings()

    def bind_mappings(self, A_padded, conv_w, conv_b):
        """Use caller-owned tensors as the mapping parameters themselves (prior training: views of one flat parameter
        vector that a fused Adam kernel updates in place): A_l as (c, round_up(c, 4)) with zero padding, conv weights
        and biases in torch's layout.  `restage_mappings()` then derives everything else from them without a copy."""
        self.A = list(A_padded)
        self.conv_w, self.conv_b = list(conv_w), list(conv_b)
        self._derived = False

    def restage_mappings(self):
        """Everything the kernels read besides A_l / conv biases, derived on the device from the current parameters:
        A_l^T, fp16 copies, upsample-folded conv weights.  No allocation after the first call."""
        self.map_generation += 1
        dev, st = self.device, stream()
        half = bool(self.tc and self.half_acts)
        if not getattr(self, "_derived", False):
            self.AT = [torch.zeros_like(a) for a in self.A]
            self.AT_h, self.A_h = [], []
            if half:   # fp16 copies for the reparameterisation GEMMs (K padded to whole 16-byte groups)
                self.AT_h = [torch.zeros(c, _round_up(c, 8), dtype=torch.float16, device=dev) for c in self.counts]
                self.A_h = [torch.zeros(c, _round_up(c, 8), dtype=torch.float16, device=dev) for c in self.counts]
            self.w_eff, self.w_eff_t, self.w_eff_k = [None] * 3, [None] * 3, [None] * 3
            for i, g in enumerate(self.geoms):
                if i == 0 and self.dense1:
                    rows = g.h * g.w * g.ic
                    cols = g.h * g.fy * g.w * g.fx * g.oc          # dense fold: 1-D / 2-D grids only
                    self.M1 = torch.zeros(rows, cols, device=dev)        # rcb_fold_dense writes the structural non-zeros only
                    self.M1T = torch.zeros(cols, rows, device=dev)
                    continue
                taps = (1 if g.kz == 1 else 2) * (1 if g.ky == 1 else 2) * (1 if g.kx == 1 else 2)
                n = g.fz * g.fy * g.fx * taps * g.ic * g.oc
                self.w_eff[i] = torch.empty(n, device=dev)
                self.w_eff_t[i] = torch.empty(n, device=dev)
                if self.tc_conv:
                    self.w_eff_k[i] = torch.empty(n, device=dev)
            g3 = self.geoms[2]
            self.f2_half = bool(self.tc_conv and self.tc and self.data_dim == 2 and g3.d == 1 and g3.fy == 2 and g3.fx == 2
                                and g3.ky == 3 and g3.kx == 3 and g3.ic == 64 and g3.oc == 16 and g3.h >= 16
                                and self.geoms[1].ic == 64)
            self._half_staged = bool(self.f2_half and self.half_acts)      # fp16 weight copies exist (half_acts at staging time)
            if self._half_staged:
                self.w3_kh = torch.empty(self.w_eff_k[2].numel(), dtype=torch.float16, device=dev)
                self.w2_kh = torch.empty(self.w_eff_k[1].numel(), dtype=torch.float16, device=dev)
                if self.dense1:
                    self.M1T_h = torch.empty(self.M1T.shape, dtype=torch.float16, device=dev)
            if self.f2_half:
                self.w3_bk = torch.empty(self.w_eff[2].numel(), device=dev)        # resident-weight data gradient
            # fp16 resident-weight data gradient of the middle stage (its incoming gradient then travels as scaled fp16)
            self.b2w = bool(self._half_staged and self._b2w_enabled
                            and self.lib.rcb_upconv_bwd_f2w_eligible(C.byref(self.geoms[1])) == 1)
            if self.b2w:
                self.w2_bk = torch.empty(self.w_eff[1].numel(), device=dev)
                self.w2_bk_h = torch.empty(self.w_eff[1].numel(), dtype=torch.float16, device=dev)
                self.w3_bk_h = torch.empty(self.w_eff[2].numel(), dtype=torch.float16, device=dev)
                if self.dense1 and self._half_da1:
                    self.M1_h = torch.empty(self.M1.shape, dtype=torch.float16, device=dev)     # conv1 data gradient on fp16 operands
            self._derived = True
        for l, c in enumerate(self.counts):
            ld = self.A[l].shape[1]
            check(self.lib.rcb_transpose(ptr(self.A[l]), ld, ptr(self.AT[l]), ld, c, c, st), "rcb_transpose")
            if half and self.AT_h:
                # (tiny, once per compression: torch copies of the two fp16 operand forms)
                self.AT_h[l][:, :c].copy_(self.AT[l][:, :c])
                self.A_h[l][:, :c].copy_(self.A[l][:, :c])
        for i, g in enumerate(self.geoms):
            if i == 0 and self.dense1:
                check(self.lib.rcb_fold_dense(ptr(self.conv_w[0]), C.byref(g), ptr(self.M1), ptr(self.M1T), st), "rcb_fold_dense")
                continue
            check(self.lib.rcb_fold_poly(ptr(self.conv_w[i]), C.byref(g), ptr(self.w_eff[i]), ptr(self.w_eff_t[i]), st), "rcb_fold_poly")
            if self.tc_conv:
                check(self.lib.rcb_fold_poly_k(ptr(self.conv_w[i]), C.byref(g), ptr(self.w_eff_k[i]), st), "rcb_fold_poly_k")
        g3 = self.geoms[2]
        if self._half_staged:
            check(self.lib.rcb_to_half(ptr(self.w_eff_k[2]), ptr(self.w3_kh), self.w3_kh.numel(), st), "rcb_to_half")
            check(self.lib.rcb_to_half(ptr(self.w_eff_k[1]), ptr(self.w2_kh), self.w2_kh.numel(), st), "rcb_to_half")
            if self.dense1:
                check(self.lib.rcb_to_half(ptr(self.M1T), ptr(self.M1T_h), self.M1T_h.numel(), st), "rcb_to_half")
        if self.f2_half:
            check(self.lib.rcb_fold_poly_bwd_f2(ptr(self.w_eff[2]), C.byref(g3), ptr(self.w3_bk), st), "rcb_fold_poly_bwd_f2")
        if self.b2w:
            check(self.lib.rcb_fold_poly_bwd_f2w(ptr(self.w_eff[1]), C.byref(self.geoms[1]), ptr(self.w2_bk), st),
                  "rcb_fold_poly_bwd_f2w")
            check(self.lib.rcb_to_half(ptr(self.w2_bk), ptr(self.w2_bk_h), self.w2_bk_h.numel(), st), "rcb_to_half")
            check(self.lib.rcb_to_half(ptr(self.w3_bk), ptr(self.w3_bk_h), self.w3_bk_h.numel(), st), "rcb_to_half")
            if getattr(self, "M1_h", None) is not None:
                check(self.lib.rcb_to_half(ptr(self.M1), ptr(self.M1_h), self.M1_h.numel(), st), "rcb_to_half")

    # -------------------------------------------------------------- workspaces --
    def workspace(self, rows: int, S: int) -> Dict[str, torch.Tensor]:
        key = (rows, S)
        ws = self._ws.get(key)
        if ws is None:
            if rows % self.R:
                raise KernelError(f"{rows} rows is not a multiple of the {self.R} patches per datum")
            items = rows * S                      # (row, sample) items: INR weights, MLP
            citems = (rows // self.R) * S         # (datum, sample) items: the stitched upsampler
            Lt = self.sp_total * self.latent_dim
            dev = self.device
            g1, g2, g3 = self.geoms
            n1 = g1.d * g1.fz * g1.h * g1.fy * g1.w * g1.fx * g1.oc
            n2 = g2.d * g2.fz * g2.h * g2.fy * g2.w * g2.fx * g2.oc
            e = lambda *s: torch.empty(*s, device=dev)
            ws = dict(hw=torch.zeros(items, self.ldw, device=dev), wt=torch.zeros(items, self.ldw, device=dev),
                      lpe=e(citems, Lt), a1=e(citems, n1), a2=e(citems, n2), pe=e(citems, self.pix_total, 16),
                      d_pe=e(citems, self.pix_total, 16), d_a2=e(citems, n2), d_a1=e(citems, n1), d_lpe=e(citems, Lt),
                      d_wt=torch.zeros(items, self.ldw, device=dev), d_hw=torch.zeros(items, self.ldw, device=dev),
                      sqerr=torch.zeros(items, device=dev), y_pred=e(items, self.pix, self.out),
                      kl=torch.zeros(1, dtype=torch.float64, device=dev), citems=citems, pe_base=None)
            if self.patch_nums is not None:
                n = np.arange(rows)
                base = ((n // self.R)[:, None] * S + np.arange(S)[None, :]) * self.pix_total + self._patch_origin[n % self.R][:, None]
                ws["pe_base"] = torch.as_tensor(base.reshape(-1).astype(np.int64), device=dev)
            self._ws[key] = ws
        return ws

    # ------------------------------------------------------ generated Fourier inputs --
    def fourier_table(self):
        """Per-axis Fourier feature table, generated on the device from coordinate indices (rcb_fourier_table), and the
        canonical (pix, n_f) input built from it.  X is identical for every datapoint of a modality (the loaders compute
        it from the pixel grid alone, data/image.py:24-27), so the tensor-core MLP regenerates it from the pixel index
        instead of reading an input tensor whenever the caller's x IS that canonical input."""
        if getattr(self, "_xtab", None) is None:
            d = self.data_dim
            nf = self.n_f // (2 * d)
            if 2 * d * nf != self.n_f or nf > 8:
                self._xtab = False
                return None
            freq = torch.exp(torch.linspace(0, float(np.log(1024)), nf))             # data/image.py:25, fp32 on the host
            fa = (C.c_float * nf)(*[float(v) for v in freq])
            offs, total = [], 0
            for sz in self.pixel_sizes:
                offs.append(total)
                total += _round_up(sz * 2 * nf, 4)
            tab = torch.zeros(total, device=self.device)
            for sz, off in zip(self.pixel_sizes, offs):
                check(self.lib.rcb_fourier_table(tab.data_ptr() + 4 * off, sz, fa, nf, stream()), "rcb_fourier_table")
            idx = torch.arange(self.pix, device=self.device)
            cos_parts, sin_parts = [], []
            rem = idx
            per_axis = []
            for sz in reversed(self.pixel_sizes):
                per_axis.append(rem % sz)
                rem = rem // sz
            per_axis.reverse()
            for sz, off, ia in zip(self.pixel_sizes, offs, per_axis):
                rows = tab[off:off + sz * 2 * nf].view(sz, 2 * nf)[ia]
                cos_parts.append(rows[:, :nf])
                sin_parts.append(rows[:, nf:])
            canon = torch.cat(cos_parts + sin_parts, 1).contiguous()                  # (pix, n_f)
            self._xtab = dict(tab=tab, offs=offs, nf=nf, canon=canon, canon_t=canon.t().contiguous())
        return self._xtab or None

    def x_is_canonical(self, x_row: torch.Tensor) -> bool:
        """Does one row of a caller's x (pix, n_f) equal the canonical Fourier input of this modality's grid?  (Host and
        device sin/cos differ in the last ulp: 2e-6 absolute on values in [-1, 1].)"""
        ft = self.fourier_table()
        if ft is None or tuple(x_row.shape) != (self.pix, self.n_f):
            return False
        return bool((x_row.to(self.device, torch.float32) - ft["canon"]).abs().max().item() <= 2e-6)

    def prepare_x(self, x):
        """(rows, pix, F) Fourier inputs -> transposed (F, pix) [shared] or (rows, F, pix).  x = None, or an x whose
        rows all equal the canonical input of the grid, selects the generated form (`self.x_generated`): the tensor-core
        MLP then takes the per-axis table and never reads X."""
        if x is None:
            ft = self.fourier_table()
            if ft is None:
                raise KernelError("this modality's inputs cannot be generated in-kernel: pass x")
            self.x_generated = True
            return ft["canon_t"], 0
        key = (x.data_ptr(), tuple(x.shape), x._version)
        if self._x_cache is not None and self._x_cache[0] == key:
            self.x_generated = self._x_cache[3]
            return self._x_cache[1], self._x_cache[2]
        if x.shape[1] != self.pix or x.shape[2] != self.n_f:
            raise KernelError(f"x has shape {tuple(x.shape)}, expected (rows, {self.pix}, {self.n_f})")
        x = x.to(device=self.device, dtype=torch.float32)
        if x.shape[0] == 1 or x.stride(0) == 0:          # broadcast view: one x for every row
            shared = True
        else:
            shared = bool((x == x[:1]).all().item())
        generated = bool(shared and self.gen_x and self.x_is_canonical(x[0]))
        # the transposed copy lives in one buffer per layout, so that captured fit steps (which hold its
        # address) keep seeing the current x
        bkey = (shared, tuple(x.shape[1:]) if shared else tuple(x.shape))
        xt = self._xt_bufs.get(bkey)
        if xt is None:
            shape = (self.n_f, self.pix) if shared else (x.shape[0], self.n_f, self.pix)
            xt = self._xt_bufs[bkey] = torch.empty(shape, device=self.device)
        if shared:
            xt.copy_(x[0].t())
            stride = 0
        else:
            xt.copy_(x.transpose(1, 2))
            stride = self.n_f * self.pix
        self._x_cache = (key, xt, stride, generated)
        self.x_generated = generated
        return xt, stride

    # ------------------------------------------------------------------ forward --
    def _sample(self, lv: LevelState, ws, S: int, noise: Noise, rows: int):
        """Level 0 writes hw and the (stitched) latent grid; levels 1, 2 add their weight
        samples to hw through the row expansion (utils.py:142-191)."""
        a = SampleArgs()
        a.loc, a.log_scale, a.mask, a.sample = ptr(lv.loc.data), ptr(lv.log_scale.data), ptr(lv.mask), ptr(lv.sample)
        a.g2p, a.perm, a.row_map = ptr(lv.g2p), ptr(lv.perm), ptr(lv.row_map)
        a.p2g = ptr(lv.p2g)
        a.fast_math = int(self.tc and self._fast_update)
        a.eps_w = ptr(noise.eps_for(lv.level))
        a.eps_l = ptr(noise.eps_l) if lv.level == 0 else None
        a.hw = ptr(ws["hw"])
        if self.half_hw:
            # single-level modalities: the weight samples are only read by the forward reparameterisation GEMM,
            # as an MMA operand -- written as fp16 (rows padded to whole 16-byte groups, padding stays zero)
            if "hw_h" not in ws:
                ws["hw_h"] = torch.zeros(rows * S, self.ldh, dtype=torch.float16, device=self.device)
            a.hw, a.hw_h = None, ptr(ws["hw_h"])
        a.lpe = ptr(ws["lpe"]) if lv.level == 0 else None
        if lv.level == 0 and self.half_acts and self.f2_half and self._half_staged and self.dense1:
            # the dense first stage reads the latent grid as an fp16 MMA operand: write it that way
            if "lpe_h" not in ws:
                ws["lpe_h"] = torch.empty(ws["lpe"].shape, dtype=torch.float16, device=self.device)
            a.lpe, a.lpe_h = None, ptr(ws["lpe_h"])
        a.lpe_slot = ptr(self.lpe_slot)
        a.rows_per_datum, a.sp_total, a.lpe_c = self.R, self.sp_total, self.latent_dim
        # generated noise is kept for the gradient kernel: 4 B/element of HBM traffic is far
        # cheaper than a second Philox + Box-Muller per element
        store = self._eps_store(ws, rows, S)
        a.eps_w_store = ptr(store[lv.level]) if a.eps_w is None else None
        a.eps_l_store = ptr(store[3]) if (lv.level == 0 and a.eps_l is None) else None
        a.seed, a.row_offset = noise.seed, noise.row_offset
        a.rows, a.S, a.P, a.n_w, a.ld_hw = rows, S, lv.P, self.W, (self.ldh if self.half_hw else self.ldw)
        a.n_l = self.L if lv.level == 0 else 0
        a.step, a.tensor_id, a.accumulate = noise.step, lv.level, int(lv.level > 0)
        a.dyn = ptr(self.step_state)
        check(self.lib.rcb_fit_sample(C.byref(a), stream()), "rcb_fit_sample")

    def _gemm(self, A, a_off, lda, B, ldb, Cm, c_off, ldc, M, N, K, bias=None, bias_mod=1, act=0, trans_a=0, acc=0,
              b_tensor=None, b_off=0, Bt=None):
        """C = A @ B.  With `Bt` (= B transposed, [N,K] row-major) and the tf32 mode the
        tcgen05 kernel is used; otherwise the fp32 SIMT engine."""
        pa = A.data_ptr() + 4 * a_off
        pc = Cm.data_ptr() + 4 * c_off
        if self.tc and Bt is not None and not trans_a:
            check(self.lib.rcb_gemm_tc(pa, lda, ptr(Bt), Bt.shape[1], pc, ldc, M, N, K, ptr(bias), bias_mod, act, acc,
                                       stream()), "rcb_gemm_tc")
            return
        pb = ptr(B) if b_tensor is None else b_tensor.data_ptr() + 4 * b_off
        check(self.lib.rcb_gemm(pa, lda, pb, ldb, pc, ldc, M, N, K, ptr(bias), bias_mod, act, trans_a, acc, stream()),
              "rcb_gemm")

    @staticmethod
    def _batch_args(a_ptrs, lda, Bt, c_ptrs, ldc, M, Ns, Ks, in_half, out_scale=1.0):
        """Argument tuple of rcb_gemm_tc_batch (host arrays kept alive next to it)."""
        nb = len(a_ptrs)
        A = (C.c_void_p * nb)(*a_ptrs)
        B = (C.c_void_p * nb)(*[t.data_ptr() for t in Bt])
        ldb = (C.c_int * nb)(*[t.shape[1] for t in Bt])
        Cp = (C.c_void_p * nb)(*c_ptrs)
        N = (C.c_int * nb)(*Ns)
        K = (C.c_int * nb)(*Ks)
        args = (nb, C.addressof(A), lda, C.addressof(B), C.addressof(ldb), C.addressof(Cp), ldc, M, C.addressof(N),
                C.addressof(K), in_half, out_scale)
        return args, (A, B, ldb, Cp, N, K, list(Bt))

    def _eps_store(self, ws, rows, S):
        st = ws.get("eps_store")
        if st is None:
            dev = self.device
            n_levels = 3 if self.patch_nums is not None else 1
            st = [torch.empty(rows, S, self.W, device=dev) if l < n_levels else None for l in range(3)]
            st.append(torch.empty(S, rows, self.L, device=dev))
            ws["eps_store"] = st
        return st

    def _upconv_fwd(self, i, src, out, citems, act):
        g = self.geoms[i]
        if self.tc_conv:
            check(self.lib.rcb_upconv_fwd_tc(ptr(src), ptr(self.w_eff_k[i]), ptr(self.conv_b[i]), ptr(out), C.byref(g),
                                             citems, act, stream()), f"rcb_upconv_fwd_tc[{i + 1}]")
        else:
            check(self.lib.rcb_upconv_fwd(ptr(src), ptr(self.w_eff[i]), ptr(self.conv_b[i]), ptr(out), C.byref(g),
                                          citems, act, stream()), f"rcb_upconv_fwd[{i + 1}]")

    def _upconv_bwd(self, i, d_out, src_act, d_src, citems):
        g = self.geoms[i]
        if self.tc_conv:
            check(self.lib.rcb_upconv_bwd_tc(ptr(d_out), ptr(self.w_eff[i]), ptr(src_act), ptr(d_src), C.byref(g), citems,
                                             stream()), f"rcb_upconv_bwd_tc[{i + 1}]")
        else:
            check(self.lib.rcb_upconv_bwd(ptr(d_out), ptr(self.w_eff_t[i]), ptr(src_act), ptr(d_src), C.byref(g), citems,
                                          stream()), f"rcb_upconv_bwd[{i + 1}]")

    def forward_features(self, lv, S: int, noise: Noise):
        """sample -> per-item INR weights (wt) and positional encodings (pe).  `lv` is the
        level-1 state, or [level1, level2, level3] for the patch modalities."""
        if self.A is None:
            raise KernelError("set_mappings() has not been called")
        levels = lv if isinstance(lv, (list, tuple)) else [lv]
        rows = levels[0].rows
        ws = self.workspace(rows, S)
        ws["red_ready"] = False                 # per-row sample sums of an earlier backward are stale now
        items = rows * S
        citems = ws["citems"]
        st = stream()
        with self.section("sample"):
            for l in levels:
                self._sample(l, ws, S, noise, rows)

        def reparam():
            with self.section("reparam_fwd"):
                if self.tc and self.batch_gemm in ("1", "fwd"):
                    half = self.half_hw
                    key = ("rp_fwd", half, self.map_generation)
                    if key not in ws:           # the four per-layer products in one launch
                        src, es, ld = (ws["hw_h"], 2, self.ldh) if half else (ws["hw"], 4, self.ldw)
                        Bt = self.AT_h if half else self.AT
                        ws[key] = self._batch_args([src.data_ptr() + es * o for o in self.offsets], ld, Bt,
                                                   [ws["wt"].data_ptr() + 4 * o for o in self.offsets], self.ldw, items,
                                                   self.counts, [_round_up(c, 8) if half else c for c in self.counts], int(half))
                    check(self.lib.rcb_gemm_tc_batch(*ws[key][0], stream()), "rcb_gemm_tc_batch")
                    return
                for l, c in enumerate(self.counts):
                    if self.half_hw:
                        check(self.lib.rcb_gemm_tc_h(ws["hw_h"].data_ptr() + 2 * self.offsets[l], self.ldh, ptr(self.AT_h[l]),
                                                     self.AT_h[l].shape[1], ws["wt"].data_ptr() + 4 * self.offsets[l], self.ldw,
                                                     items, c, _round_up(c, 8), None, 1, 0, 0, stream()), "rcb_gemm_tc_h")
                        continue
                    self._gemm(ws["hw"], self.offsets[l], self.ldw, self.A[l], self.A[l].shape[1],
                               ws["wt"], self.offsets[l], self.ldw, items, c, c, Bt=self.AT[l])
        join = self._fork(reparam)
        self._upsample_chain(ws, citems)
        join()
        return ws

    def _upsample_chain(self, ws, citems: int):
        """lpe -> a1 -> a2 -> pe: the three folded nearest-up + conv stages (prior_model.py:47-59)."""
        g1, g2, g3 = self.geoms
        half = ws["a2_is_half"] = bool(self.half_acts and self.f2_half and self._half_staged)
        if half and "a2h" not in ws:
            ws["a1h"] = torch.empty(ws["a1"].shape, dtype=torch.float16, device=self.device)
            ws["a2h"] = torch.empty(ws["a2"].shape, dtype=torch.float16, device=self.device)
        with self.section("conv1_fwd"):
            if self.dense1:
                Lt = self.M1.shape[0]
                if half:
                    check(self.lib.rcb_gemm_tc_hh(ptr(ws["lpe_h"]), Lt, ptr(self.M1T_h), self.M1T.shape[1], ptr(ws["a1h"]),
                                                  ws["a1"].shape[1], citems, self.M1.shape[1], Lt, ptr(self.conv_b[0]), g1.oc, 1,
                                                  stream()), "rcb_gemm_tc_hh")
                else:
                    self._gemm(ws["lpe"], 0, Lt, self.M1, self.M1.shape[1], ws["a1"], 0, ws["a1"].shape[1],
                               citems, self.M1.shape[1], Lt, bias=self.conv_b[0], bias_mod=g1.oc, act=1, Bt=self.M1T)
            elif half:
                check(self.lib.rcb_upconv_fwd_tc_oh(ptr(ws["lpe"]), ptr(self.w_eff_k[0]), ptr(self.conv_b[0]), ptr(ws["a1h"]),
                                                    C.byref(g1), citems, 1, stream()), "rcb_upconv_fwd_tc_oh[1]")
            else:
                self._upconv_fwd(0, ws["lpe"], ws["a1"], citems, 1)
        with self.section("conv2_fwd"):
            if half:
                check(self.lib.rcb_upconv_fwd_tc_hh(ptr(ws["a1h"]), ptr(self.w2_kh), ptr(self.conv_b[1]), ptr(ws["a2h"]),
                                                    C.byref(g2), citems, 1, stream()), "rcb_upconv_fwd_tc_hh[2]")
            else:
                self._upconv_fwd(1, ws["a1"], ws["a2"], citems, 1)
        # the MLP rounds the positional encodings to fp16 for its first MMA anyway: stored that way when both ends
        # are the fp16 kernels
        half_pe = ws["pe_is_half"] = bool(half and self.half_pe and self.tc_mlp and self.n_f == 16)
        if half_pe and "pe_h" not in ws:
            ws["pe_h"] = torch.empty(ws["pe"].shape, dtype=torch.float16, device=self.device)
        with self.section("conv3_fwd"):
            if half_pe:
                check(self.lib.rcb_upconv_fwd_tc_hh(ptr(ws["a2h"]), ptr(self.w3_kh), ptr(self.conv_b[2]), ptr(ws["pe_h"]),
                                                    C.byref(g3), citems, 0, stream()), "rcb_upconv_fwd_tc_hh[3]")
            elif half:
                check(self.lib.rcb_upconv_fwd_tc_h(ptr(ws["a2h"]), ptr(self.w3_kh), ptr(self.conv_b[2]), ptr(ws["pe"]),
                                                   C.byref(g3), citems, 0, stream()), "rcb_upconv_fwd_tc_h[3]")
            else:
                self._upconv_fwd(2, ws["a2"], ws["pe"], citems, 0)

    def upsample_latents(self, latent: torch.Tensor) -> torch.Tensor:
        """(S, rows, L) latent grids (channel-last per row) -> (rows, S, pixels, 16) positional encodings: the kernel
        form of utils.map_lpe_to_inr_inputs (utils.py:4-120), stitching the rows of a datum into one grid, running the
        folded upsampler on it and cutting the result back into patches.  Evaluation only (no autograd graph)."""
        if self.A is None and getattr(self, "conv_w", None) is None:
            raise KernelError("the upsampler weights have not been staged (set_mappings / set_upsampler)")
        S, rows = int(latent.shape[0]), int(latent.shape[1])
        ws = self.workspace(rows, S)
        D, C, sp = rows // self.R, self.latent_dim, self.sp
        lat = latent.to(self.device, torch.float32).reshape(S, D, self.R, sp, C).permute(1, 0, 2, 3, 4)     # (D, S, R, sp, C)
        dst = ws["lpe"].view(D, S, self.sp_total, C)
        if self.lpe_slot is None:
            dst.copy_(lat.reshape(D, S, self.sp_total, C))
        else:
            dst[:, :, self.lpe_slot.reshape(-1).long(), :] = lat.reshape(D, S, self.R * sp, C)
        half = bool(self.half_acts and self.f2_half and self._half_staged)
        if half and self.dense1:
            if "lpe_h" not in ws:
                ws["lpe_h"] = torch.empty(ws["lpe"].shape, dtype=torch.float16, device=self.device)
            ws["lpe_h"].copy_(ws["lpe"])
        self._upsample_chain(ws, ws["citems"])
        pe = (ws["pe_h"] if ws.get("pe_is_half") else ws["pe"]).float().view(D, S, self.pix_total, 16)
        if self.patch_nums is None:
            return pe.view(rows, S, self.pix, 16).clone()
        if getattr(self, "_patch_pixels", None) is None:          # stitched-grid offset of every pixel of every patch
            pad = 3 - self.data_dim
            fp = [1] * pad + list(self.full_pixels)
            pp = [1] * pad + list(self.pixel_sizes)
            z, y, x = np.meshgrid(np.arange(pp[0]), np.arange(pp[1]), np.arange(pp[2]), indexing="ij")
            inner = (z * fp[1] * fp[2] + y * fp[2] + x).reshape(-1)
            self._patch_pixels = torch.as_tensor((self._patch_origin[:, None] + inner[None, :]).reshape(-1), device=self.device)
        out = pe[:, :, self._patch_pixels, :].view(D, S, self.R, self.pix, 16)
        return out.permute(0, 2, 1, 3, 4).reshape(rows, S, self.pix, 16).contiguous()

    def set_step_state(self, buf: torch.Tensor, seed: int, step: int, adam: dict, t: int, beta_scalar: float = -1.0):
        """Write the per-step scalars a captured step reads from device memory (rcb_step_state)."""
        check(self.lib.rcb_set_step_state(ptr(buf), seed, step, adam["lr"] / (1.0 - adam["b1"] ** t),
                                          math.sqrt(1.0 - adam["b2"] ** t), beta_scalar, stream()), "rcb_set_step_state")

    def _fork(self, fn):
        """Run fn on the side stream, ordered after everything queued so far on the current stream;
        returns the join that makes the current stream wait for it."""
        if not self.overlap:
            fn()
            return lambda: None
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
            self._ev_fork, self._ev_join = torch.cuda.Event(), torch.cuda.Event()
        self._ev_fork.record(main)
        with torch.cuda.stream(self._side):
            self._side.wait_event(self._ev_fork)
            fn()
            self._ev_join.record(self._side)
        return lambda: main.wait_event(self._ev_join)

    def mlp(self, ws, rows: int, S: int, x, mode: int, y=None, dy=None, coef: float = 0.0):
        xt, stride = self.prepare_x(x)
        a = MlpArgs()
        a.wt, a.xt, a.pe = ptr(ws["wt"]), ptr(xt), ptr(ws["pe"])
        if ws.get("pe_is_half"):
            a.pe, a.pe_half = ptr(ws["pe_h"]), 1
        a.y, a.dy, a.y_pred = ptr(y), ptr(dy), ptr(ws["y_pred"])
        a.d_pe, a.d_wt, a.sqerr = ptr(ws["d_pe"]), ptr(ws["d_wt"]), ptr(ws["sqerr"])
        a.x_row_stride = stride
        a.pe_base = ptr(ws["pe_base"])
        a.pitch_z, a.pitch_y, a.ph, a.pw = self.pitch_z, self.pitch_y, self.ph, self.pw
        a.items, a.S, a.pix, a.n_f, a.out, a.ld_w, a.mode = rows * S, S, self.pix, self.n_f, self.out, self.ldw, mode
        use_tc = self.tc_mlp and self.n_f in (16, 18)
        if use_tc and self.x_generated:
            ft = self.fourier_table()
            a.x_tab, a.x_axes, a.x_nfreq = ptr(ft["tab"]), self.data_dim, ft["nf"]
            for i in range(self.data_dim):
                a.x_size[i], a.x_off[i] = self.pixel_sizes[i], ft["offs"][i]
        if use_tc and mode == 2 and coef == 0.0:
            # the tensor-core kernel keeps its weight-gradient operands in fp16: hand it a power-of-two
            # scale that brings the caller's dy to O(1) (host sync; this is the autograd path, not the fit loop)
            amax = float(dy.abs().max())
            coef = 2.0 ** round(-math.log2(amax)) if amax > 0.0 and math.isfinite(amax) else 1.0
        a.coef, a.w0 = coef, self.w0
        # power of two that brings d_pe (and what the upsampler's adjoint makes of it) to O(1): mode 1 writes
        # coef * (residual chain), mode 2 (chain of dy * coef) / coef
        ws["bwd_scale"] = 0.0
        ws["d_pe_is_half"] = False
        if use_tc and coef > 0.0 and mode in (1, 2):
            ws["bwd_scale"] = 2.0 ** round(-math.log2(coef)) if mode == 1 else coef
            if self.b2w and self._half_dpe and ws.get("a2_is_half"):
                # d pe leaves the kernel as fp16 in the chain's own units (true / coef resp. true * coef): the fp16
                # data gradients of the two x2 stages carry that unit, the last one multiplies it out
                if "d_pe_h" not in ws:
                    ws["d_pe_h"] = torch.empty(ws["d_pe"].shape, dtype=torch.float16, device=self.device)
                a.d_pe_h = ptr(ws["d_pe_h"])
                ws["bwd_scale"] = 1.0 / coef if mode == 1 else coef
                ws["d_pe_is_half"] = True
        ws["d_wt_is_half"] = False
        if use_tc and mode == 1 and self.half_hw and self.half_dwt:
            # the weight gradients are only read by the data-gradient reparameterisation GEMM: written as fp16
            # (scaled by 2^8, clamped), the GEMM divides the scale out again
            if "d_wt_h" not in ws:
                ws["d_wt_h"] = torch.zeros(rows * S, self.ldh, dtype=torch.float16, device=self.device)
            a.d_wt_h, a.ld_wh, a.d_wt_h_scale = ptr(ws["d_wt_h"]), self.ldh, 256.0
            ws["d_wt_is_half"] = True
        with self.section("mlp_fwd" if mode == 0 else "mlp_fwd_bwd"):
            if use_tc:
                check(self.lib.rcb_mlp_tc(C.byref(a), stream()), "rcb_mlp_tc")
            else:
                check(self.lib.rcb_mlp(C.byref(a), stream()), "rcb_mlp")

    # ----------------------------------------------------------------- backward --
    def backward_features(self, ws, rows: int, S: int):
        """d_pe, d_wt -> d_lpe, d_hw (data gradients only; the mappings are frozen)."""
        items = rows * S
        citems = ws["citems"]
        st = stream()
        g1, g2, g3 = self.geoms

        def reparam():
            with self.section("reparam_bwd"):
                if ws.get("d_wt_is_half"):
                    key = ("rp_bwd_h", self.map_generation)
                    if key not in ws:
                        ws[key] = self._batch_args([ws["d_wt_h"].data_ptr() + 2 * o for o in self.offsets], self.ldh, self.A_h,
                                                   [ws["d_hw"].data_ptr() + 4 * o for o in self.offsets], self.ldw, items,
                                                   self.counts, [_round_up(c, 8) for c in self.counts], 1, 1.0 / 256.0)
                    check(self.lib.rcb_gemm_tc_batch(*ws[key][0], stream()), "rcb_gemm_tc_batch")
                    return
                if self.tc and self.batch_gemm == "1":
                    key = ("rp_bwd", self.map_generation)
                    if key not in ws:
                        ws[key] = self._batch_args([ws["d_wt"].data_ptr() + 4 * o for o in self.offsets], self.ldw, self.A,
                                                   [ws["d_hw"].data_ptr() + 4 * o for o in self.offsets], self.ldw, items,
                                                   self.counts, self.counts, 0)
                    check(self.lib.rcb_gemm_tc_batch(*ws[key][0], stream()), "rcb_gemm_tc_batch")
                    return
                for l, c in enumerate(self.counts):
                    self._gemm(ws["d_wt"], self.offsets[l], self.ldw, self.AT[l], self.AT[l].shape[1],
                               ws["d_hw"], self.offsets[l], self.ldw, items, c, c, Bt=self.A[l])
        join = self._fork(reparam)
        b2w = bool(self.b2w and ws.get("a2_is_half") and ws.get("bwd_scale", 0.0) > 0.0)
        if b2w:
            # fp16 gradient between the two x2 stages, in units of 1 / bwd_scale
            if "d_a2_h" not in ws:
                ws["d_a2_h"] = torch.empty(ws["d_a2"].shape, dtype=torch.float16, device=self.device)
            sc = float(ws["bwd_scale"])
            with self.section("conv3_bwd"):
                if ws.get("d_pe_is_half"):
                    check(self.lib.rcb_upconv_bwd_f2_hh(ptr(ws["d_pe_h"]), ptr(self.w3_bk_h), ptr(ws["a2h"]), 2, ptr(ws["d_a2_h"]),
                                                        1.0, C.byref(g3), citems, stream()), "rcb_upconv_bwd_f2_hh[3]")
                else:
                    check(self.lib.rcb_upconv_bwd_f2_oh(ptr(ws["d_pe"]), ptr(self.w3_bk), ptr(ws["a2h"]), 2, ptr(ws["d_a2_h"]), sc,
                                                        C.byref(g3), citems, stream()), "rcb_upconv_bwd_f2_oh[3]")
            # d_a1 stays fp16 (same unit) when the stage below is the dense fp16 GEMM, which then multiplies the unit out
            da1_half = bool(self.dense1 and self.M1_h is not None and ws.get("d_pe_is_half"))
            with self.section("conv2_bwd"):
                if da1_half:
                    if "d_a1_h" not in ws:
                        ws["d_a1_h"] = torch.empty(ws["d_a1"].shape, dtype=torch.float16, device=self.device)
                    check(self.lib.rcb_upconv_bwd_f2w_oh(ptr(ws["d_a2_h"]), ptr(self.w2_bk_h), ptr(ws["a1h"]), ptr(ws["d_a1_h"]), 1.0,
                                                         C.byref(g2), citems, stream()), "rcb_upconv_bwd_f2w_oh[2]")
                else:
                    check(self.lib.rcb_upconv_bwd_f2w(ptr(ws["d_a2_h"]), ptr(self.w2_bk_h), ptr(ws["a1h"]), ptr(ws["d_a1"]), 1.0 / sc,
                                                      C.byref(g2), citems, stream()), "rcb_upconv_bwd_f2w[2]")
        else:
            with self.section("conv3_bwd"):
                if self.f2_half:
                    half = bool(ws.get("a2_is_half"))
                    check(self.lib.rcb_upconv_bwd_f2(ptr(ws["d_pe"]), ptr(self.w3_bk), ptr(ws["a2h"] if half else ws["a2"]),
                                                     2 if half else 1, ptr(ws["d_a2"]), C.byref(g3), citems, stream()),
                          "rcb_upconv_bwd_f2[3]")
                else:
                    self._upconv_bwd(2, ws["d_pe"], ws["a2"], ws["d_a2"], citems)
            with self.section("conv2_bwd"):
                if ws.get("a2_is_half"):
                    check(self.lib.rcb_upconv_bwd_tc_ah(ptr(ws["d_a2"]), ptr(self.w_eff[1]), ptr(ws["a1h"]), ptr(ws["d_a1"]),
                                                        C.byref(g2), citems, stream()), "rcb_upconv_bwd_tc_ah[2]")
                else:
                    self._upconv_bwd(1, ws["d_a2"], ws["a1"], ws["d_a1"], citems)
        with self.section("conv1_bwd"):
            if b2w and da1_half:
                Lt, n1 = self.M1.shape
                key = ("c1_bwd_h", self.map_generation, sc)
                if key not in ws:
                    ws[key] = self._batch_args([ws["d_a1_h"].data_ptr()], n1, [self.M1_h], [ws["d_lpe"].data_ptr()], Lt, citems,
                                               [Lt], [n1], 1, 1.0 / sc)
                check(self.lib.rcb_gemm_tc_batch(*ws[key][0], stream()), "rcb_gemm_tc_batch[conv1_bwd]")
            elif self.dense1:
                Lt = self.M1.shape[0]
                self._gemm(ws["d_a1"], 0, ws["d_a1"].shape[1], self.M1T, self.M1T.shape[1], ws["d_lpe"], 0, Lt,
                           citems, Lt, self.M1T.shape[0], Bt=self.M1)
            else:
                self._upconv_bwd(0, ws["d_a1"], None, ws["d_lpe"], citems)
        join()

    def backward_mappings(self, ws, rows: int, S: int, out=None):
        """Gradients of the learned mappings (prior training): dA_l = hw_l^T d_wt_l, and the
        upsampler's conv weights/biases through the adjoints of the folds.  Call after
        backward_features (it consumes d_pe, d_a2, d_a1, d_wt).  `out` = {"A": [...], "conv1.weight": ..., ...}
        gives the destination tensors (views of one flat gradient vector); otherwise they live in the workspace."""
        if ws.get("a2_is_half"):
            raise KernelError("the mapping gradients read the upsampler activations in fp32: set engine.half_acts = False")
        items = rows * S
        citems = ws["citems"]
        st = stream()
        dev = self.device
        g = ws.setdefault("map_grads", {})
        if not g or g.get("_out") is not out:
            keep = {k: v for k, v in g.items() if k.startswith("eff") or k == "dM1"}
            g.clear()
            g.update(keep)
            g["_out"] = out
            if out is not None:
                g.update({k: v for k, v in out.items()})
            else:
                g["A"] = [torch.zeros(c, _round_up(c, 4), device=dev) for c in self.counts]
            for i, geo in enumerate(self.geoms):
                kshape = (geo.oc, geo.ic) + (geo.kz, geo.ky, geo.kx)[3 - self.data_dim:]
                if out is None:
                    g[f"conv{i + 1}.weight"] = torch.zeros(*kshape, device=dev)
                    g[f"conv{i + 1}.bias"] = torch.zeros(geo.oc, device=dev)
                taps = (1 if geo.kz == 1 else 2) * (1 if geo.ky == 1 else 2) * (1 if geo.kx == 1 else 2)
                if f"eff{i}" not in g:
                    g[f"eff{i}"] = torch.zeros(geo.fz * geo.fy * geo.fx * taps * geo.ic * geo.oc, device=dev)
            if self.dense1 and "dM1" not in g:
                g["dM1"] = torch.zeros_like(self.M1)
        with self.section("reparam_wgrad"):
            if self.tc:
                # dA_l = hw_l^T d_wt_l with K = items: transpose both once (K-major), then the tcgen05 GEMM
                ldt = _round_up(items, 4)
                if "hwT" not in ws or ws["hwT"].shape[1] != ldt:
                    ws["hwT"] = torch.zeros(self.ldw, ldt, device=dev)
                    ws["dwtT"] = torch.zeros(self.ldw, ldt, device=dev)
                check(self.lib.rcb_transpose(ptr(ws["hw"]), self.ldw, ptr(ws["hwT"]), ldt, items, self.W, st), "rcb_transpose")
                check(self.lib.rcb_transpose(ptr(ws["d_wt"]), self.ldw, ptr(ws["dwtT"]), ldt, items, self.W, st), "rcb_transpose")
                for l, c in enumerate(self.counts):
                    pa = ws["hwT"].data_ptr() + 4 * self.offsets[l] * ldt
                    pb = ws["dwtT"].data_ptr() + 4 * self.offsets[l] * ldt
                    check(self.lib.rcb_gemm_tc(pa, ldt, pb, ldt, ptr(g["A"][l]), g["A"][l].shape[1], c, c, items,
                                               None, 1, 0, 0, st), "rcb_gemm_tc")
            else:
                for l, c in enumerate(self.counts):
                    self._gemm(ws["hw"], self.offsets[l], self.ldw, None, self.ldw, g["A"][l], 0, g["A"][l].shape[1],
                               c, c, items, trans_a=1, b_tensor=ws["d_wt"], b_off=self.offsets[l])
        srcs = [ws["lpe"], ws["a1"], ws["a2"]]
        douts = [ws["d_a1"], ws["d_a2"], ws["d_pe"]]
        with self.section("conv_wgrad"):
            for i, geo in enumerate(self.geoms):
                out_px = geo.d * geo.fz * geo.h * geo.fy * geo.w * geo.fx
                check(self.lib.rcb_colsum(ptr(douts[i]), citems * out_px, geo.oc, geo.oc, ptr(g[f"conv{i + 1}.bias"]), st),
                      "rcb_colsum")
                if i == 0 and self.dense1:
                    Lt = self.M1.shape[0]
                    n1 = g["dM1"].shape[1]
                    if self.tc and n1 % 4 == 0:
                        # dM1 = lpe^T d_a1 with K = items: K-major copies, then the tcgen05 GEMM
                        ldt = _round_up(citems, 4)
                        if "lpeT" not in ws or ws["lpeT"].shape[1] != ldt:
                            ws["lpeT"] = torch.zeros(Lt, ldt, device=dev)
                            ws["da1T"] = torch.zeros(n1, ldt, device=dev)
                        check(self.lib.rcb_transpose(ptr(ws["lpe"]), Lt, ptr(ws["lpeT"]), ldt, citems, Lt, st), "rcb_transpose")
                        check(self.lib.rcb_transpose(ptr(ws["d_a1"]), n1, ptr(ws["da1T"]), ldt, citems, n1, st), "rcb_transpose")
                        check(self.lib.rcb_gemm_tc(ptr(ws["lpeT"]), ldt, ptr(ws["da1T"]), ldt, ptr(g["dM1"]), n1, Lt, n1, citems,
                                                   None, 1, 0, 0, st), "rcb_gemm_tc")
                    else:
                        self._gemm(ws["lpe"], 0, Lt, None, n1, g["dM1"], 0, n1,
                                   Lt, n1, citems, trans_a=1, b_tensor=ws["d_a1"], b_off=0)
                    check(self.lib.rcb_unfold_dense(ptr(g["dM1"]), C.byref(geo), ptr(g["conv1.weight"]), st),
                          "rcb_unfold_dense")
                    continue
                in_px = geo.d * geo.h * geo.w
                if (self.tc_conv and self.data_dim == 2 and geo.ic in (64, 128) and geo.oc % 16 == 0
                        and (geo.h * geo.w) % 32 == 0 and geo.w % 4 == 0):
                    # tcgen05 weight gradient on channel-major (K-major) copies of the activations
                    kt = f"wgT{i}"
                    if kt not in ws or ws[kt][0].shape[1] != citems * in_px:
                        ws[kt] = (torch.empty(3 * geo.ic, citems * in_px, device=dev), torch.empty(geo.oc, citems * out_px, device=dev))
                    srcT, doutT = ws[kt]
                    check(self.lib.rcb_transpose_xshift(ptr(srcs[i]), ptr(srcT), citems * in_px, geo.ic, geo.w, st),
                          "rcb_transpose_xshift")
                    check(self.lib.rcb_transpose_phases(ptr(douts[i]), ptr(doutT), citems * out_px, geo.oc, geo.h, geo.w,
                                                        geo.fy, geo.fx, st), "rcb_transpose_phases")
                    check(self.lib.rcb_upconv_wgrad_tc(ptr(srcT), ptr(doutT), ptr(g[f"eff{i}"]), C.byref(geo), citems, st),
                          "rcb_upconv_wgrad_tc")
                else:
                    check(self.lib.rcb_upconv_wgrad(ptr(srcs[i]), ptr(douts[i]), ptr(g[f"eff{i}"]), C.byref(geo), citems, st),
                          "rcb_upconv_wgrad")
                check(self.lib.rcb_unfold_poly(ptr(g[f"eff{i}"]), C.byref(geo), ptr(g[f"conv{i + 1}.weight"]), st),
                      "rcb_unfold_poly")
        return g

    def reduce_samples(self, levels, ws, S: int, noise: Noise, rows: int):
        """Patch modalities: sum the data gradient (and gradient x noise of every level) over the S samples once per
        patch row (rcb_fit_reduce); the per-level update kernels then read these sums.  No-op for the single-level
        modalities, whose row kernel reduces in shared memory."""
        if self.patch_nums is None:
            return
        red = ws.get("red")
        if red is None:
            e = lambda n: torch.empty(rows, n, device=self.device)
            red = ws["red"] = dict(mu=e(self.W), sig=[e(self.W) for _ in levels], mu_l=e(self.L), sig_l=e(self.L))
        store = ws.get("eps_store")
        a = ReduceArgs()
        a.d_hw, a.d_lpe, a.lpe_slot = ptr(ws["d_hw"]), ptr(ws["d_lpe"]), ptr(self.lpe_slot)
        for l, lv in enumerate(levels):
            eps = noise.eps_for(lv.level)
            if eps is None:
                if store is None:
                    raise KernelError("reduce_samples needs the noise kept by the sampling kernel")
                eps = store[lv.level]
            a.eps_w[l] = ptr(eps)
            a.red_sig[l] = ptr(red["sig"][l])
        a.eps_l = ptr(noise.eps_l if noise.eps_l is not None else store[3])
        a.red_mu, a.red_mu_l, a.red_sig_l = ptr(red["mu"]), ptr(red["mu_l"]), ptr(red["sig_l"])
        a.rows, a.S, a.n_w, a.n_l, a.ld_hw, a.n_levels = rows, S, self.W, self.L, self.ldw, len(levels)
        a.rows_per_datum, a.sp_total, a.lpe_c = self.R, self.sp_total, self.latent_dim
        with self.section("update"):
            check(self.lib.rcb_fit_reduce(C.byref(a), stream()), "rcb_fit_reduce")
        ws["red_ready"] = True

    def update(self, lv: LevelState, ws, S: int, noise: Noise, *, with_data_grads: bool, adam: Optional[dict],
               g_loc=None, g_log_scale=None, grad_scale: float = 1.0, kl_out: Optional[torch.Tensor] = None,
               rows: Optional[int] = None):
        a = UpdateArgs()
        a.loc, a.log_scale, a.mask = ptr(lv.loc.data), ptr(lv.log_scale.data), ptr(lv.mask)
        a.p_loc, a.p_log_scale, a.beta, a.group_idx = ptr(lv.p_loc), ptr(lv.p_log_scale), ptr(lv.beta), ptr(lv.group_idx)
        a.p2g, a.perm_inv, a.row_children = ptr(lv.p2g), ptr(lv.perm_inv), ptr(lv.row_children)
        a.d_hw = ptr(ws["d_hw"]) if with_data_grads else None
        a.d_lpe = ptr(ws["d_lpe"]) if (with_data_grads and lv.level == 0) else None
        a.eps_w = ptr(noise.eps_for(lv.level))
        a.eps_l = ptr(noise.eps_l) if lv.level == 0 else None
        if with_data_grads and ws is not None and "eps_store" in ws:      # noise kept by the sampling kernel
            if a.eps_w is None:
                a.eps_w = ptr(ws["eps_store"][lv.level])
            if a.eps_l is None and lv.level == 0:
                a.eps_l = ptr(ws["eps_store"][3])
        a.lpe_slot = ptr(self.lpe_slot)
        a.rows_per_datum, a.sp_total, a.lpe_c = self.R, self.sp_total, self.latent_dim
        if with_data_grads and ws is not None and ws.get("red_ready"):
            red = ws["red"]
            a.red_mu, a.red_sig = ptr(red["mu"]), ptr(red["sig"][lv.level])
            if lv.level == 0:
                a.red_mu_l, a.red_sig_l = ptr(red["mu_l"]), ptr(red["sig_l"])
        a.g_loc, a.g_log_scale = ptr(g_loc), ptr(g_log_scale)
        a.kl_out = ptr(kl_out)
        a.seed, a.row_offset = noise.seed, noise.row_offset
        a.src_rows, a.rows, a.n_children, a.S, a.P = lv.rows, (rows if rows is not None else lv.rows), lv.n_children, S, lv.P
        a.n_w, a.n_l, a.ld_hw, a.G = self.W, (self.L if lv.level == 0 else 0), self.ldw, lv.G
        a.step, a.tensor_id = noise.step, lv.level
        a.beta_scalar, a.grad_scale = float(lv.beta_scalar), grad_scale
        a.fast_math = int(self.tc and self._fast_update)      # never in the fp32 parity configuration
        a.p_scale_direct = int(lv.p_scale_direct)
        a.dyn = ptr(self.step_state)
        if adam is not None:
            st_ = lv.adam
            if self.step_state is None:      # a captured step gets t from set_step_state(); its caller counts
                st_["t"] += 1
            t = max(st_["t"], 1)
            a.adam = 1
            a.m1_loc, a.v_loc, a.m1_ls, a.v_ls = ptr(st_["m1_loc"]), ptr(st_["v_loc"]), ptr(st_["m1_ls"]), ptr(st_["v_ls"])
            a.b1, a.b2, a.adam_eps = adam["b1"], adam["b2"], adam["eps"]
            a.adam_step_size = adam["lr"] / (1.0 - adam["b1"] ** t)
            a.adam_bc2_sqrt = math.sqrt(1.0 - adam["b2"] ** t)
        else:
            a.adam = 0
        with self.section("update"):
            check(self.lib.rcb_fit_update(C.byref(a), stream()), "rcb_fit_update")

    # --------------------------------------------------------------- block KL ----
    def group_kl(self, lv: LevelState) -> torch.Tensor:
        check(self.lib.rcb_group_kl(ptr(lv.loc.data), ptr(lv.log_scale.data), ptr(lv.p_loc), ptr(lv.p_log_scale),
                                    ptr(lv.group_start), ptr(lv.group_end), ptr(lv.group_kl), lv.rows, lv.P, lv.G,
                                    stream()), "rcb_group_kl")
        return lv.group_kl

    def anneal(self, lv: LevelState, step: float, upper: float, lower: float, bits: float):
        check(self.lib.rcb_anneal_beta(ptr(lv.beta), ptr(lv.group_kl), ptr(lv.coded), lv.rows, lv.G,
                                       step, upper, lower, bits, stream()), "rcb_anneal_beta")
