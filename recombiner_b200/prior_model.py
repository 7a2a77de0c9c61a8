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
        """(B, 128, *grid) -> (B, 16, *pixels) as the reference module computes it (prior_model.py:47-59), on the folded
        polyphase kernels.  Evaluation only: the fit path differentiates through FitEngine, not through this call."""
        from . import utils as _utils
        from ._lib import KernelError as _KE
        if not x.is_cuda:
            raise _KE("Upsample.forward runs on the sm_100a kernels: pass a CUDA tensor (no CPU fallback)")
        d = self.kernel_dim
        grid = list(x.shape[2:])
        factors = [1] * d
        for i in (1, 2, 3):
            f = getattr(self, f"up{i}").scale_factor
            f = [int(v) for v in f] if isinstance(f, (tuple, list)) else [int(f)] * d
            factors = [a * b for a, b in zip(factors, f)]
        pixels = [g * f for g, f in zip(grid, factors)]
        eng = _utils._engine_for(self, int(x.shape[1]), pixels, factors, False, None, d, x.device)
        lat = x.detach().movedim(1, -1).reshape(1, x.shape[0], -1)            # (S = 1, rows, L) channel-last
        pe = eng.upsample_latents(lat)                                        # (rows, 1, pixels, 16)
        return pe[:, 0].reshape(x.shape[0], *pixels, pe.shape[-1]).movedim(-1, 1).contiguous()


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


# --------------------------------------------------------------------------- #
# prior-training model (prior_model.py:62-262) on the sm_100a kernels
# --------------------------------------------------------------------------- #
import math as _math

import torch.nn.functional as _F

from . import _lib as _rcb_lib
from ._lib import KernelError, check as _check, ptr as _ptr, stream as _stream
from .utils import count_net_params


class SharedMappings:
    """The learned mappings of prior training (A_l of LinearTransform, conv weights / biases of Upsample) as ONE flat
    fp32 parameter vector in the layout the kernels read, with a gradient vector and Adam moments of the same layout.

    The weight-gradient kernels write straight into `grad` (no per-tensor copies), data-parallel ranks exchange it with
    a single all-reduce, and `rcb_adam_flat` steps `theta` in place -- the reference's `Adam(list(A) + list(upsample))`
    (prior_model.py:224-227) without per-step staging.  A_l is stored padded to (c, round_up(c, 4)) (16-byte rows for
    the tensor-core GEMMs); the padding columns have zero gradient and stay zero."""

    def __init__(self, engine, linear_transform, upsample_net):
        dev = engine.device
        self.lt, self.up = linear_transform, upsample_net
        self.segments = []                     # (name, parameter, offset, stored shape, logical column count)
        off = 0
        for l, (p, c) in enumerate(zip(linear_transform.A, engine.counts)):
            ld = (c + 3) // 4 * 4
            self.segments.append((f"A{l}", p, off, (c, ld), c))
            off += c * ld
        for name, p in upsample_net.named_parameters():
            off = (off + 3) // 4 * 4
            self.segments.append((name, p, off, tuple(p.shape), None))
            off += p.numel()
        self.n = (off + 3) // 4 * 4
        self.theta, self.grad, self.m, self.v = (torch.zeros(self.n, device=dev) for _ in range(4))
        self.t = 0
        view = lambda buf, o, shape: buf[o:o + int(np.prod(shape))].view(shape)
        self.params = {nm: view(self.theta, o, shp) for nm, _, o, shp, _ in self.segments}
        self.grads = {nm: view(self.grad, o, shp) for nm, _, o, shp, _ in self.segments}
        n_a = len(engine.counts)
        self.grad_out = {"A": [self.grads[f"A{l}"] for l in range(n_a)]}
        self.grad_out.update({nm: self.grads[nm] for nm, _, _, _, c in self.segments if c is None})
        engine.bind_mappings([self.params[f"A{l}"] for l in range(n_a)],
                             [self.params[f"conv{i}.weight"] for i in (1, 2, 3)],
                             [self.params[f"conv{i}.bias"] for i in (1, 2, 3)])

    def load(self):
        """modules -> theta; fresh Adam moments (the reference re-creates its optimiser on every train() call)."""
        with torch.no_grad():
            for nm, p, _, _, c in self.segments:
                dst = self.params[nm]
                (dst[:, :c] if c is not None else dst).copy_(p.detach())
        self.m.zero_(); self.v.zero_()
        self.t = 0

    def store(self):
        """theta -> modules (so checkpoints pickle the trained values)."""
        with torch.no_grad():
            for nm, p, _, _, c in self.segments:
                src = self.params[nm]
                p.copy_(src[:, :c] if c is not None else src)


class PriorBNNmodel(nn.Module):
    """All training rows' factorised Gaussian posteriors, fitted jointly with the shared
    mappings (LinearTransform, Upsample) against the current prior.

    Same constructor, attributes (`loc`, `log_scale`, `lpe_loc`, `lpe_log_scale`, `st`,
    `dims`) and methods (`forward`, `calculate_kl`, `train`) as the reference class.  The
    weights and the latent grid of a row live in one (rows, W+L) tensor in parameter order
    so one kernel pass covers both; `loc` / `lpe_loc` are views of it.  Rows may be a shard
    of the training set: with torch.distributed initialised, `train` all-reduces the
    shared-mapping gradients every step."""

    def __init__(self, in_dim, hidden_dims, out_dim, train_size, data_dim, pixel_sizes, upsample_factors, latent_dim,
                 patch, patch_nums, hierarchical_patch_nums, random_seed=42, device="cuda", init_log_scale=-4, c=6.,
                 w0=30., layer_scales=None, paddings=None, row_offset=0, global_train_size=None, precision=None):
        super().__init__()
        from .engine import FitEngine, LevelState
        dev = torch.device(device)
        if dev.type != "cuda":
            raise KernelError("recombiner_b200.PriorBNNmodel runs on CUDA (sm_100a) only -- no CPU fallback")
        _rcb_lib.use_device(dev)
        dev = torch.device("cuda", torch.cuda.current_device()) if dev.index is None else dev
        self.random_seed, self.device = random_seed, dev
        self.n_layers = len(hidden_dims) + 1
        self.dims = [in_dim] + list(hidden_dims) + [out_dim]
        self.patch = patch
        self.act = lambda v: torch.sin(w0 * v)
        self.st = lambda v: _F.softplus(v, beta=1, threshold=20) / 6
        self.data_dim, self.train_size, self.latent_dim = data_dim, train_size, latent_dim
        self.pixel_sizes, self.upsample_factors = pixel_sizes, upsample_factors
        self.patch_nums, self.hierarchical_patch_nums = patch_nums, hierarchical_patch_nums
        self.net_params_list, self.cum_param_sizes = count_net_params(in_dim, hidden_dims, out_dim)
        self.row_offset = int(row_offset)
        W = int(self.cum_param_sizes[-1])
        self._lpe_shape = [pixel_sizes[i] // upsample_factors[i] for i in range(data_dim)] + [latent_dim]
        L = int(np.prod(self._lpe_shape))
        self._W, self._L = W, L
        # reference init order under torch.manual_seed(seed): loc (rand), [h_loc, hh_loc (rand)],
        # then lpe_loc (randn) (prior_model.py:100-110).  A shard draws the global tensors and
        # keeps its rows.
        n_glob = int(global_train_size) if global_train_size is not None else train_size
        R = int(np.prod(patch_nums)) if patch else 1
        n2 = int(np.prod(hierarchical_patch_nums['level2'])) if patch else 1
        n3 = int(np.prod(hierarchical_patch_nums['level3'])) if patch else 1
        if train_size % R or self.row_offset % R or (patch and (train_size % n2 or train_size % n3)):
            raise KernelError("patch modalities shard by whole datum: rows must be a multiple of the patch count")
        torch.manual_seed(random_seed)
        w_std = np.sqrt(c / hidden_dims[-1]) / w0
        loc = torch.rand(n_glob, W) * w_std * 2 - w_std
        if patch:
            h_loc = torch.rand(n_glob // n2, W) * w_std * 2 - w_std
            hh_loc = torch.rand(n_glob // n3, W) * w_std * 2 - w_std
        lpe = torch.randn(n_glob, *self._lpe_shape) * 0.1
        rows = slice(self.row_offset, self.row_offset + train_size)
        both = torch.cat([loc[rows], lpe[rows].reshape(train_size, L)], 1)
        self._loc_all = nn.Parameter(both.to(dev).contiguous())
        self._log_scale_all = nn.Parameter(torch.zeros(train_size, W + L, device=dev) + init_log_scale)
        self.engine = FitEngine(self.dims, data_dim, pixel_sizes, upsample_factors, latent_dim,
                                layer_scales if layer_scales is not None else [4, 2, 2],
                                paddings if paddings is not None else [2, 1, 1], w0, dev, precision=precision,
                                patch_nums=patch_nums if patch else None)
        self.engine.half_acts = False         # the weight gradients of the upsampler read its activations in fp32
        zero = torch.zeros(W + L)
        self._lv = LevelState(self._loc_all, self._log_scale_all, zero, zero, None, None, None, None, None, 0.0, dev)
        self._lv.p_scale_direct = True
        self._levels = [self._lv]
        if patch:
            def shard(t, div):
                return t[self.row_offset // div:(self.row_offset + train_size) // div]
            self.h_loc = nn.Parameter(shard(h_loc, n2).to(dev).contiguous())
            self.h_log_scale = nn.Parameter(torch.zeros(train_size // n2, W, device=dev) + init_log_scale)
            self.hh_loc = nn.Parameter(shard(hh_loc, n3).to(dev).contiguous())
            self.hh_log_scale = nn.Parameter(torch.zeros(train_size // n3, W, device=dev) + init_log_scale)
            l2 = hierarchical_patch_nums['level2']
            ng = [patch_nums[i] // l2[i] for i in range(data_dim)]
            n = np.arange(train_size)
            pc = np.unravel_index(n % R, patch_nums)
            grp = np.ravel_multi_index([pc[i] // l2[i] for i in range(data_dim)], ng)
            maps = [(n // R) * int(np.prod(ng)) + grp, n // R]
            zw = torch.zeros(W)
            for li, (lo, ls) in enumerate(((self.h_loc, self.h_log_scale), (self.hh_loc, self.hh_log_scale)), start=1):
                lv = LevelState(lo, ls, zw, zw, None, None, None, None, None, 0.0, dev)
                lv.p_scale_direct, lv.level = True, li
                lv.set_expansion(maps[li - 1])
                self._levels.append(lv)
        self._call = 0

    def global_rows(self, level: int = 0) -> int:
        """Rows of a level over all ranks (cached: the sharding does not change during training)."""
        cache = self.__dict__.setdefault("_global_rows", {})
        if level not in cache:
            cache[level] = global_count(self._levels[level].rows, self.device)
        return cache[level]

    # views with the reference's attribute names
    @property
    def loc(self):
        return self._loc_all[:, :self._W]

    @property
    def log_scale(self):
        return self._log_scale_all[:, :self._W]

    @property
    def lpe_loc(self):
        return self._loc_all[:, self._W:].reshape(self.train_size, *self._lpe_shape)

    @property
    def lpe_log_scale(self):
        return self._log_scale_all[:, self._W:].reshape(self.train_size, *self._lpe_shape)

    def group_to_layer(self, params, layer_idx):
        lo = 0 if layer_idx == 0 else self.cum_param_sizes[layer_idx - 1]
        return params[..., lo:self.cum_param_sizes[layer_idx]]

    def layer_to_weight(self, in_dim, out_dim, layer_param):
        return layer_param[:, out_dim:].reshape(-1, in_dim, out_dim), layer_param[:, :out_dim]

    # ------------------------------------------------------------------ pieces --
    def _noise(self, eps=None, step=0):
        from .engine import Noise
        if eps is not None:
            dev = self.device
            return Noise(eps_w=eps["w"].to(dev).contiguous(),
                         eps_l=eps["lpe"].to(dev).reshape(1, self.train_size, self._L).contiguous(),
                         eps_h=eps["h"].to(dev).contiguous() if "h" in eps else None,
                         eps_hh=eps["hh"].to(dev).contiguous() if "hh" in eps else None)
        base = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
        return Noise(seed=((int(self.random_seed) & 0x7fffffff) << 32) | base, step=step, row_offset=self.row_offset)

    def _set_prior(self, prior_loc, prior_scale, prior_lpe_loc, prior_lpe_scale, prior_h_loc=None, prior_h_scale=None,
                   prior_hh_loc=None, prior_hh_scale=None):
        """Copy the prior into the levels' persistent buffers (captured steps hold their addresses)."""
        f = lambda t: t.reshape(-1).to(self.device, torch.float32)
        W = self._W

        def put(lv, loc_parts, scale_parts):
            n = sum(p.numel() for p in loc_parts)
            if lv.p_loc.numel() != n or not getattr(lv, "_prior_owned", False):
                lv.p_loc, lv.p_log_scale = torch.empty(n, device=self.device), torch.empty(n, device=self.device)
                lv._prior_owned = True
            o = 0
            for pl, ps in zip(loc_parts, scale_parts):
                lv.p_loc[o:o + pl.numel()].copy_(f(pl))
                lv.p_log_scale[o:o + ps.numel()].copy_(f(ps))
                o += pl.numel()
        put(self._lv, [prior_loc, prior_lpe_loc], [prior_scale, prior_lpe_scale])
        if self.patch:
            put(self._levels[1], [prior_h_loc], [prior_h_scale])
            put(self._levels[2], [prior_hh_loc], [prior_hh_scale])

    def forward(self, x, linear_transform, upsample_net, gradient_through_A=True, eps=None):
        """Single-sample reconstruction (rows, pixels, out) (prior_model.py:129-179).  Evaluation
        only: gradients are produced by `train` / `loss_and_grads`, not by autograd."""
        assert x.shape[0] == self.train_size
        eng = self.engine
        eng.set_mappings(list(linear_transform.A), upsample_net.state_dict())
        ws = eng.forward_features(self._levels, 1, self._noise(eps))
        eng.mlp(ws, self.train_size, 1, x.to(self.device), mode=0)
        return ws["y_pred"].view(self.train_size, eng.pix, eng.out).clone()

    def _kl(self, beta: float):
        """(sum of beta*KL as f64 device scalar, d/dloc, d/dlog_scale) of the current posterior."""
        from .engine import Noise
        kl = torch.zeros(1, dtype=torch.float64, device=self.device)
        grads = []
        for lv in self._levels:
            g_loc, g_ls = torch.empty_like(lv.loc.data), torch.empty_like(lv.log_scale.data)
            lv.beta_scalar = beta
            self.engine.update(lv, None, 1, Noise(), with_data_grads=False, adam=None, g_loc=g_loc, g_log_scale=g_ls,
                               kl_out=kl, rows=self.train_size)
            grads.append((g_loc, g_ls))
        return kl, grads

    def calculate_kl(self, prior_loc, prior_scale, prior_lpe_loc, prior_lpe_scale, prior_h_loc=None, prior_h_scale=None,
                     prior_hh_loc=None, prior_hh_scale=None):
        """sum KL(q || p) over weights and latent grid (prior_model.py:181-200)."""
        self._set_prior(prior_loc, prior_scale, prior_lpe_loc, prior_lpe_scale, prior_h_loc, prior_h_scale,
                        prior_hh_loc, prior_hh_scale)
        return self._kl(1.0)[0].to(torch.float32).reshape(())

    def loss_and_grads(self, x, y, priors, linear_transform, upsample_net, kl_beta, eps=None, training_mappings=True):
        """One forward/backward without an optimiser step (parity tests): returns
        (mse*N, sum KL, gradient dict) of loss = N*mean((y_hat-y)^2) + kl_beta * sum KL."""
        eng, lv, N = self.engine, self._lv, self.train_size
        self._set_prior(*priors)
        eng.set_mappings(list(linear_transform.A), upsample_net.state_dict())
        noise = self._noise(eps)
        ws = eng.forward_features(self._levels, 1, noise)
        eng.mlp(ws, N, 1, x.to(self.device), mode=1, y=y.to(self.device).contiguous(), coef=2.0 / (eng.pix * eng.out))
        eng.backward_features(ws, N, 1)
        grads = {}
        if training_mappings:
            g = eng.backward_mappings(ws, N, 1)
            for l, c in enumerate(eng.counts):
                grads[f"A{l}"] = g["A"][l][:, :c].clone()
            for k in ("conv1", "conv2", "conv3"):
                grads[k + ".weight"], grads[k + ".bias"] = g[k + ".weight"].clone(), g[k + ".bias"].clone()
        kl = torch.zeros(1, dtype=torch.float64, device=self.device)
        per_level = []
        eng.reduce_samples(self._levels, ws, 1, noise, N)
        for l in self._levels:
            g_loc, g_ls = torch.empty_like(l.loc.data), torch.empty_like(l.log_scale.data)
            l.beta_scalar = float(kl_beta)
            eng.update(l, ws, 1, noise, with_data_grads=True, adam=None, g_loc=g_loc, g_log_scale=g_ls, kl_out=kl, rows=N)
            per_level.append((g_loc, g_ls))
        W = self._W
        g_loc, g_ls = per_level[0]
        grads.update(loc=g_loc[:, :W], log_scale=g_ls[:, :W], lpe_loc=g_loc[:, W:].reshape(N, *self._lpe_shape),
                     lpe_log_scale=g_ls[:, W:].reshape(N, *self._lpe_shape))
        if self.patch:
            grads.update(h_loc=per_level[1][0], h_log_scale=per_level[1][1], hh_loc=per_level[2][0],
                         hh_log_scale=per_level[2][1])
        mse = ws["sqerr"].sum() / (eng.pix * eng.out)
        return mse, kl / max(float(kl_beta), 1e-300), grads

    def _shared(self, linear_transform, upsample_net) -> SharedMappings:
        sm = self.__dict__.get("_shared_mappings")
        if sm is None or sm.lt is not linear_transform or sm.up is not upsample_net:
            sm = self._shared_mappings = SharedMappings(self.engine, linear_transform, upsample_net)
            self._step_graphs = {}
        return sm

    def _step_body(self, sm, x, y, noise, cfg, coef, training_mappings, kl_step, world):
        """One full-batch step on the current stream: re-derive the staged mappings from theta, forward, loss,
        backward, mapping gradients into sm.grad, [all-reduce], posterior KL-gradient + Adam, mapping Adam."""
        import torch.distributed as dist
        eng, N = self.engine, self.train_size
        eng.restage_mappings()
        ws = eng.forward_features(self._levels, 1, noise)
        eng.mlp(ws, N, 1, x, mode=1, y=y, coef=coef)
        eng.backward_features(ws, N, 1)
        work = None
        if training_mappings:
            eng.backward_mappings(ws, N, 1, out=sm.grad_out)
            if world > 1:        # ONE collective for all shared-mapping gradients, overlapped with the posterior update
                work = dist.all_reduce(sm.grad, op=dist.ReduceOp.SUM, async_op=True)
        kl_step.zero_()
        eng.reduce_samples(self._levels, ws, 1, noise, N)
        for l in self._levels:
            eng.update(l, ws, 1, noise, with_data_grads=True, adam=cfg, kl_out=kl_step, rows=N)
        if training_mappings:
            if work is not None:
                work.wait()
            t = max(self._levels[0].adam["t"], 1)
            _check(eng.lib.rcb_adam_flat(_ptr(sm.theta), _ptr(sm.grad), _ptr(sm.m), _ptr(sm.v), sm.n,
                                         cfg["lr"] / (1.0 - cfg["b1"] ** t), _math.sqrt(1.0 - cfg["b2"] ** t),
                                         cfg["b1"], cfg["b2"], cfg["eps"], _ptr(eng.step_state), _stream()), "rcb_adam_flat")
        return ws

    def train(self, n_epoch=True, lr=2e-4, x=None, y=None, prior_loc=None, prior_scale=None, prior_lpe_loc=None,
              prior_lpe_scale=None, prior_h_loc=None, prior_h_scale=None, prior_hh_loc=None, prior_hh_scale=None,
              linear_transform=None, upsample_net=None, kl_beta=1e-8, training_mappings=True, verbose=False):
        """n_epoch full-batch Adam steps on loss = N*mean((y_hat-y)^2) + kl_beta*sum KL over the
        posteriors and (optionally) the mappings (prior_model.py:202-262).  Returns
        (last mse / N, KL / N, list of per-step ELBOs).

        Device-resident: the mappings live in one flat parameter vector (SharedMappings), nothing is allocated per
        step, the per-step loss terms stay on the device (one read-back per call), and on a single GPU the whole step
        is replayed from a captured CUDA graph with the noise key, Adam bias corrections and kl_beta in device memory
        (rcb_step_state).  With several ranks the step is launched kernel by kernel (no NCCL call inside a captured
        graph) and the single all-reduce of the flat gradient vector runs beside the posterior update."""
        if isinstance(n_epoch, bool):           # nn.Module.train(mode) / .eval()
            return super().train(n_epoch)
        import ctypes as C
        import os
        import torch.distributed as dist
        from ._lib import StepState
        eng, N = self.engine, self.train_size
        x = x.to(self.device)
        y = y.to(self.device, torch.float32).contiguous()
        self._set_prior(prior_loc, prior_scale, prior_lpe_loc, prior_lpe_scale, prior_h_loc, prior_h_scale,
                        prior_hh_loc, prior_hh_scale)
        for l in self._levels:
            l.beta_scalar = float(kl_beta)
            l.reset_adam()                       # the reference re-creates Adam on every call (:224-227)
        sm = self._shared(linear_transform, upsample_net)
        sm.load()
        cfg = dict(lr=float(lr), b1=0.9, b2=0.999, eps=1e-8)
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        n_total = self.global_rows(0)            # true row count over all shards (they may be uneven)
        coef = 2.0 / (eng.pix * eng.out)
        stats = torch.zeros(max(n_epoch, 1), 2, dtype=torch.float64, device=self.device)     # per step: (mse*N, beta*KL)
        kl_step = self.__dict__.setdefault("_kl_step", torch.zeros(1, dtype=torch.float64, device=self.device))
        base_noise = self._noise()
        Noise = type(base_noise)
        use_graph = (os.environ.get("RECOMBINER_GRAPH", "1") != "0" and eng.timer is None
                     and not torch.cuda.is_current_stream_capturing())
        xt, x_stride = eng.prepare_x(x)
        key = (xt.data_ptr(), x_stride, eng.x_generated, y.data_ptr(), tuple(y.shape), bool(training_mappings), world, cfg["b1"], cfg["b2"],
               cfg["eps"], self.row_offset, id(sm)) + tuple(v for l in self._levels for v in (l.p_loc.data_ptr(), l.loc.data_ptr()))
        graphs = self.__dict__.setdefault("_step_graphs", {})
        state = self.__dict__.setdefault("_step_state", torch.zeros(C.sizeof(StepState), dtype=torch.uint8, device=self.device))
        it = range(n_epoch)
        if verbose:
            from tqdm import tqdm
            it = tqdm(it)
        ws = None
        for i in it:
            noise = Noise(seed=base_noise.seed, step=i, row_offset=self.row_offset)
            entry = graphs.get(key) if use_graph else None
            if use_graph and entry is None and key in graphs and world == 1:
                # second sight of this configuration: capture the step (single GPU: all of it)
                graph = torch.cuda.CUDAGraph()
                eng.step_state = state
                try:
                    with torch.cuda.graph(graph):
                        ws = self._step_body(sm, x, y, noise, cfg, coef, training_mappings, kl_step, world)
                finally:
                    eng.step_state = None
                entry = graphs[key] = (graph, ws)
            if entry is not None:
                graph, ws = entry
                eng.set_step_state(state, noise.seed, noise.step, cfg, i + 1, beta_scalar=float(kl_beta))
                graph.replay()
                for l in self._levels:
                    l.adam["t"] = i + 1
            else:
                if use_graph and key not in graphs:
                    graphs[key] = None           # first sight: run eagerly (sizes workspaces, TMA maps)
                ws = self._step_body(sm, x, y, noise, cfg, coef, training_mappings, kl_step, world)
            _check(eng.lib.rcb_step_stats(_ptr(ws["sqerr"]), N, 1.0 / (eng.pix * eng.out), _ptr(kl_step),
                                          stats.data_ptr() + 16 * i, _stream()), "rcb_step_stats")
        sm.store()
        if world > 1:
            dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        kl_final = self._kl(1.0)[0]
        if world > 1:
            dist.all_reduce(kl_final, op=dist.ReduceOp.SUM)
        stats = stats.cpu()
        elbo = (-(stats[:n_epoch, 0] + stats[:n_epoch, 1])).tolist()
        mse_last = float(stats[n_epoch - 1, 0]) if n_epoch > 0 else float("nan")
        return mse_last / n_total, float(kl_final.item()) / n_total, elbo


def global_count(n_local: int, device) -> int:
    """Sum of a per-rank row count over the process group (ranks may hold uneven shards:
    parallel.shard_rows hands the remainder to the first ranks)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return int(n_local)
    t = torch.tensor([int(n_local)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def _level_prior(lib, loc, log_scale, n_rows, n_total):
    import torch.distributed as dist
    P, dev = loc.shape[1], loc.device
    stats = torch.empty(3 * P, dtype=torch.float64, device=dev)
    _check(lib.rcb_prior_suffstats(_ptr(loc), _ptr(log_scale), _ptr(stats), n_rows, P, _stream()), "rcb_prior_suffstats")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    p_loc, p_scale = torch.empty(P, device=dev), torch.empty(P, device=dev)
    _check(lib.rcb_prior_from_stats(_ptr(stats), _ptr(p_loc), _ptr(p_scale), n_total, P, _stream()), "rcb_prior_from_stats")
    return p_loc, p_scale


def em_prior_update(model: "PriorBNNmodel"):
    """Closed-form prior update from all posteriors (main_prior_training.py:157-172):
    mu_p = mean_n mu_q, sigma_p = sqrt(mean_n sigma_q^2 + var_n mu_q) (unbiased variance),
    from f64 sufficient statistics that are all-reduced across shards.
    Returns (prior_loc, prior_scale, prior_lpe_loc, prior_lpe_scale[, prior_h_loc, prior_h_scale,
    prior_hh_loc, prior_hh_scale])."""
    import torch.distributed as dist
    lib = _rcb_lib.load()
    W = model._W
    out = []
    for li, lv in enumerate(model._levels):
        p_loc, p_scale = _level_prior(lib, lv.loc.data, lv.log_scale.data, lv.rows, model.global_rows(li))
        if li == 0:
            out += [p_loc[:W], p_scale[:W], p_loc[W:].reshape(model._lpe_shape), p_scale[W:].reshape(model._lpe_shape)]
        else:
            out += [p_loc, p_scale]
    return tuple(out)
