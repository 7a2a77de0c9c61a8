"""Test-time (compression) model: per-datapoint variational INR fitting + REC.

Drop-in host mirror of the reference `test_model.TestBNNmodel`
(test_model.py:33-856): same constructor keywords, attributes and methods, but
every tensor op of the hot path runs in the sm_100a kernels of
librecombiner_b200.so (no CPU fallback, no torch autograd graph in the loop):

  predict / train        -> FitEngine (sample, reparam GEMMs, folded upsampler,
                            fused SIREN MLP fwd+loss+bwd, KL-gradient + Adam)
  update_annealing_factors -> rcb_group_kl + rcb_anneal_beta (no host round trip)
  sample_group / compress_group / compress_posteriors
                         -> rcb_rec_table + batched rcb_rec_encode (one launch codes
                            one block of *every* row; the reference loops rows in
                            Python, test_model.py:806-818)
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional

import numpy as np
import torch
from torch import nn

from . import rec as _rec
from .engine import FitEngine, LevelState, Noise
from ._lib import COUNTERS, KernelError, StepState
from .utils import count_net_params, metric


class Sine(nn.Module):
    """sin(w0 x) (kept for API compatibility; the kernels evaluate it in the MLP epilogue)."""

    def __init__(self, w0=1.):
        super().__init__()
        self.w0 = w0

    def forward(self, x):
        return torch.sin(self.w0 * x)


class _Predict(torch.autograd.Function):
    """y_pred = INR(posterior sample); backward re-runs the fused MLP in gradient mode.
    Tensor inputs: (loc, log_scale) of every level, in level order."""

    @staticmethod
    def forward(ctx, model, x, S, noise, *params):
        levels = model._levels
        rows = levels[0].rows
        ws = model.engine.forward_features(levels, S, noise)
        model.engine.mlp(ws, rows, S, x, mode=0)
        ctx.model, ctx.x, ctx.S, ctx.noise = model, x, S, noise
        model._generation += 1
        ctx.generation = model._generation
        return ws["y_pred"].view(rows, S, model.engine.pix, model.engine.out).clone()

    @staticmethod
    def backward(ctx, dy):
        model, S, noise = ctx.model, ctx.S, ctx.noise
        if ctx.generation != model._generation:
            raise KernelError("predict() workspace was overwritten by a later forward before backward()")
        levels = model._levels
        rows = levels[0].rows
        ws = model.engine.workspace(rows, S)
        dy = dy.contiguous().view(rows * S, model.engine.pix, model.engine.out)
        model.engine.mlp(ws, rows, S, ctx.x, mode=2, dy=dy)
        model.engine.backward_features(ws, rows, S)
        model.engine.reduce_samples(levels, ws, S, noise, rows)
        grads = []
        for lv in levels:
            g_loc, g_ls = torch.empty_like(lv.loc.data), torch.empty_like(lv.log_scale.data)
            saved_beta = lv.beta
            lv.beta = torch.zeros_like(saved_beta)      # data term only; KL has its own Function
            try:
                model.engine.update(lv, ws, S, noise, with_data_grads=True, adam=None, g_loc=g_loc, g_log_scale=g_ls,
                                    rows=rows)
            finally:
                lv.beta = saved_beta
            grads += [g_loc, g_ls]
        return (None, None, None, None, *grads)


class _WeightedKL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, *params):
        kl = torch.zeros(1, dtype=torch.float64, device=model.device)
        grads = []
        for lv in model._levels:
            g_loc, g_ls = torch.empty_like(lv.loc.data), torch.empty_like(lv.log_scale.data)
            model.engine.update(lv, None, 1, Noise(), with_data_grads=False, adam=None, g_loc=g_loc, g_log_scale=g_ls,
                                kl_out=kl, rows=model._levels[0].rows)
            grads += [g_loc, g_ls]
        ctx.save_for_backward(*grads)
        return kl.to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, g):
        return (None, *[g * t for t in ctx.saved_tensors])


def _level_prop(li, attr, post=None):
    """Read-only view of one level's device state under the reference's attribute name."""
    def get(self):
        v = getattr(self._levels[li], attr)
        return post(v) if post else v
    return property(get)


def _as_bool(t):
    return t.bool().cpu().numpy()


def _as_f64(t):
    return t.cpu().numpy().astype(np.float64)


class TestBNNmodel(nn.Module):
    __test__ = False   # not a pytest class
    _PREFIX = ("", "h_", "hh_")

    def __init__(self,
                 in_dim, hidden_dims, out_dim, number_of_datapoints, upsample_factors, latent_dim, data_dim,
                 pixel_sizes, patch, patch_nums, hierarchical_patch_nums, dataset,
                 linear_transform=None, upsample_net=None,
                 p_loc=None, p_log_scale=None, init_log_scale=-4., param_to_group=None, group_to_param=None,
                 n_groups=None, group_start_index=None, group_end_index=None, group_idx=None,
                 h_p_loc=None, h_p_log_scale=None, h_init_log_scale=-4., h_param_to_group=None,
                 h_group_to_param=None, h_n_groups=None, h_group_start_index=None, h_group_end_index=None,
                 h_group_idx=None,
                 hh_p_loc=None, hh_p_log_scale=None, hh_init_log_scale=-4., hh_param_to_group=None,
                 hh_group_to_param=None, hh_n_groups=None, hh_group_start_index=None, hh_group_end_index=None,
                 hh_group_idx=None,
                 w0=30., c=6., random_seed=42, device='cuda', kl_upper_buffer=0., kl_lower_buffer=0.4,
                 kl_adjust_gap=10, initial_beta=1e-8, beta_step_size=0.05, row_offset=0, layer_scales=None,
                 paddings=None, precision=None):
        super().__init__()
        dev = torch.device(device)
        if dev.type != "cuda":
            raise KernelError("recombiner_b200.TestBNNmodel runs on CUDA (sm_100a) only -- there is no CPU "
                              "fallback; the CPU restatement used for parity lives in oracle/")
        from ._lib import use_device
        use_device(dev)
        dev = torch.device("cuda", torch.cuda.current_device()) if dev.index is None else dev
        self.bit_per_group = 16
        self.n_layers = len(hidden_dims) + 1
        self.dims = [in_dim] + list(hidden_dims) + [out_dim]
        self.upsample_factors, self.latent_dim, self.data_dim = upsample_factors, latent_dim, data_dim
        self.patch, self.patch_nums, self.pixel_sizes = patch, patch_nums, pixel_sizes
        self.hierarchical_patch_nums = hierarchical_patch_nums
        self.linear_transform, self.upsample_net = linear_transform, upsample_net
        self.device, self.dataset, self.random_seed = dev, dataset, random_seed
        self.row_offset = int(row_offset)
        for m in (linear_transform, upsample_net):
            if m is not None:
                for p in m.parameters():
                    p.requires_grad = False
        _, self.cum_param_sizes = count_net_params(in_dim, hidden_dims, out_dim)
        self.beta_step_size, self.kl_upper_buffer = beta_step_size, kl_upper_buffer
        self.kl_lower_buffer, self.kl_adjust_gap = kl_lower_buffer, kl_adjust_gap

        rows = number_of_datapoints
        R = int(np.prod(patch_nums)) if patch else 1
        level_rows = [rows]
        if patch:
            level_rows += [rows // int(np.prod(hierarchical_patch_nums['level2'])),
                           rows // int(np.prod(hierarchical_patch_nums['level3']))]
        given = [
            (p_loc, p_log_scale, init_log_scale, param_to_group, group_to_param, n_groups, group_start_index,
             group_end_index, group_idx),
            (h_p_loc, h_p_log_scale, h_init_log_scale, h_param_to_group, h_group_to_param, h_n_groups,
             h_group_start_index, h_group_end_index, h_group_idx),
            (hh_p_loc, hh_p_log_scale, hh_init_log_scale, hh_param_to_group, hh_group_to_param, hh_n_groups,
             hh_group_start_index, hh_group_end_index, hh_group_idx)]
        self._levels = []
        for li, n_rows in enumerate(level_rows):
            pre = self._PREFIX[li]
            pl, pls, ils, p2g, g2p, ng, gs, ge, gi = given[li]
            P = pl.shape[0]
            ils = ils.to(dev) if torch.is_tensor(ils) else ils
            loc = nn.Parameter(pl.detach().to(dev, torch.float32)[None, :].repeat(n_rows, 1).contiguous())
            log_scale = nn.Parameter((torch.zeros(n_rows, P, device=dev) + ils).contiguous())
            setattr(self, pre + "loc", loc)
            setattr(self, pre + "log_scale", log_scale)
            setattr(self, pre + "p_loc", pl.detach().clone().to(dev, torch.float32))
            setattr(self, pre + "p_log_scale", pls.detach().clone().to(dev, torch.float32))
            for nm, v in (("param_to_group", p2g), ("group_to_param", g2p), ("n_groups", ng),
                          ("group_start_index", gs), ("group_end_index", ge), ("group_idx", gi)):
                setattr(self, pre + nm, v)
            lv = LevelState(loc, log_scale, getattr(self, pre + "p_loc"), getattr(self, pre + "p_log_scale"), gi, gs, ge,
                            g2p, p2g, initial_beta, dev)
            lv.level = li
            setattr(self, pre + "compressed_sample_std", 1e-15 + torch.zeros(n_rows, P, device=dev))
            self._levels.append(lv)
        self._lv = self._levels[0]
        if patch:
            # per-column row permutations of levels 1 and 2 (test_model.py:182-208)
            for li in (0, 1):
                lv = self._levels[li]
                perm = np.stack([np.random.RandomState(c).choice(lv.rows, lv.rows, False) for c in range(lv.P)], 1)
                np.random.seed(None)
                lv.set_permutation(perm)
                setattr(self, self._PREFIX[li] + "permute_patch_x_g2p", perm)
                setattr(self, self._PREFIX[li] + "permute_patch_x_p2g", np.argsort(perm, axis=0))
            # expansion of level-2 / level-3 rows over the patches (utils.py:151-189)
            l2 = hierarchical_patch_nums['level2']
            ng = [patch_nums[i] // l2[i] for i in range(data_dim)]
            n = np.arange(rows)
            pc = np.unravel_index(n % R, patch_nums)
            grp = np.ravel_multi_index([pc[i] // l2[i] for i in range(data_dim)], ng)
            self._levels[1].set_expansion((n // R) * int(np.prod(ng)) + grp)
            self._levels[2].set_expansion(n // R)

        cfg_scales = layer_scales if layer_scales is not None else [4, 2, 2]
        self.engine = FitEngine(self.dims, data_dim, pixel_sizes, upsample_factors, latent_dim, cfg_scales,
                                paddings if paddings is not None else [2, 1, 1], w0, dev, precision=precision,
                                patch_nums=patch_nums if patch else None)
        if linear_transform is not None and upsample_net is not None:
            self.engine.set_mappings(list(linear_transform.A), upsample_net.state_dict())
        self.act = Sine(w0)
        self.st = lambda v: torch.nn.functional.softplus(v, beta=1, threshold=20) / 6
        pixels = np.prod(pixel_sizes)
        self.bpp = (self.n_groups * self.bit_per_group) / pixels
        if patch:
            self.bpp += (self.h_n_groups * self.bit_per_group) / pixels / np.prod(hierarchical_patch_nums['level2'])
            self.bpp += (self.hh_n_groups * self.bit_per_group) / pixels / np.prod(hierarchical_patch_nums['level3'])
        if self.dataset == 'audio':
            self.bpp = self.bpp / (3 / 48000) / 1000
        print("Model Initialized. Expected bpp is %.2f" % self.bpp, flush=True)

        self.g_samples = None
        self._g_dev = None
        self.group_samples = {}
        self.h_group_samples, self.hh_group_samples = {}, {}
        self._tables = None
        self._generation = 0
        self._adam_owner = None
        self._graphs = {}
        self.use_graph = os.environ.get("RECOMBINER_GRAPH", "1") != "0"

    # ------------------------------------------------------- reference attributes --
    kl_beta = property(lambda self: self._levels[0].beta, lambda self, v: self._set_beta(0, v))
    h_kl_beta = property(lambda self: self._levels[1].beta, lambda self, v: self._set_beta(1, v))
    hh_kl_beta = property(lambda self: self._levels[2].beta, lambda self, v: self._set_beta(2, v))
    compressed_mask = _level_prop(0, "mask")
    h_compressed_mask = _level_prop(1, "mask")
    hh_compressed_mask = _level_prop(2, "mask")
    compressed_sample = _level_prop(0, "sample")
    h_compressed_sample = _level_prop(1, "sample")
    hh_compressed_sample = _level_prop(2, "sample")
    compressed_mask_groupwise = _level_prop(0, "coded", _as_bool)
    h_compressed_mask_groupwise = _level_prop(1, "coded", _as_bool)
    hh_compressed_mask_groupwise = _level_prop(2, "coded", _as_bool)
    # (rows, G) float64 numpy, as the reference stores and np.savetxt's them (test_model.py:221)
    compressed_idx_groupwise = _level_prop(0, "idx", _as_f64)
    h_compressed_idx_groupwise = _level_prop(1, "idx", _as_f64)
    hh_compressed_idx_groupwise = _level_prop(2, "idx", _as_f64)

    def _set_beta(self, li, v):
        lv = self._levels[li]
        lv.beta = torch.as_tensor(v, dtype=torch.float32).to(self.device).expand(lv.rows, lv.G).contiguous()

    def group_to_layer(self, param, layer_idx):
        lo = 0 if layer_idx == 0 else self.cum_param_sizes[layer_idx - 1]
        return param[..., lo:self.cum_param_sizes[layer_idx]]

    def layer_to_weight(self, in_dim, out_dim, layer_param):
        lead = layer_param.shape[:-1]
        bias = layer_param[..., :out_dim].unsqueeze(-2)
        weights = layer_param[..., out_dim:].reshape(*lead, in_dim, out_dim)
        return weights, bias

    def _params(self):
        out = []
        for lv in self._levels:
            out += [lv.loc, lv.log_scale]
        return out

    # -------------------------------------------------------------------- forward --
    def _noise(self, random_seed, eps=None) -> Noise:
        if eps is not None:
            dev = self.device
            return Noise(eps_w=eps["w"].to(dev).contiguous(), eps_l=eps["lpe"].to(dev).contiguous(),
                         eps_h=eps["h"].to(dev).contiguous() if "h" in eps else None,
                         eps_hh=eps["hh"].to(dev).contiguous() if "hh" in eps else None)
        if random_seed is None:
            random_seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())   # global torch RNG, like randn_like
        seed = ((int(self.random_seed) & 0x7fffffff) << 32) | (int(random_seed) & 0xffffffff)
        return Noise(seed=seed, step=0, row_offset=self.row_offset)

    def predict(self, x, random_seed=None, sample_size=1, eps=None):
        """MC forward (test_model.py:283-355).  `eps` (dict with 'lpe' (S,N,L), 'w' (N,S,W) and,
        for patch modalities, 'h' / 'hh' (N,S,W)) injects the noise for parity tests;
        otherwise it is Philox-generated in-kernel from (model seed, random_seed)."""
        noise = self._noise(random_seed, eps)
        y = _Predict.apply(self, x, sample_size, noise, *self._params())
        return y[:, 0] if sample_size == 1 else y

    def calculate_kl(self):
        """sum over levels of sum_{n,p} beta[n, g(p)] KL(q_np || p_p)  (test_model.py:357-377)."""
        return _WeightedKL.apply(self, *self._params())

    def update_annealing_factors(self, update=True):
        """Per-(row, block) KL in nats; optionally anneal beta (test_model.py:379-439).
        Returns (rows, G) float64 numpy -- one array per level for patch modalities."""
        out = [k.cpu().numpy() for k in self._annealing(update)]
        return tuple(out) if self.patch else out[0]

    def _annealing(self, update: bool):
        kls = []
        for lv in self._levels:
            kls.append(self.engine.group_kl(lv))
            if update:
                self.engine.anneal(lv, self.beta_step_size, self.kl_upper_buffer, self.kl_lower_buffer,
                                   float(self.bit_per_group))
        return kls

    # ------------------------------------------------------------------------ REC --
    def get_gumbel_sample(self):
        n = int(np.ceil(2 ** self.bit_per_group))
        g = _rec.gumbel_sequence(self.random_seed, n)
        self.g_samples = torch.from_numpy(g)
        self._g_dev = self.g_samples.to(self.device)

    def _ensure_rec(self, n_cand: int):
        if self.g_samples is None or self._g_dev is None or self._g_dev.numel() < n_cand:
            self.get_gumbel_sample()
        if self._tables is None or self._tables.n != n_cand:
            self._tables = _rec.CandidateTables(self.random_seed, n_cand, self.device)
            for lv in self._levels:       # all levels share tables of equal block size and the Gumbel noise
                sizes = lv.group_end_host - lv.group_start_host
                lv.tables_ptr = self._tables.pointer_array(sizes)
                lv.max_D = int(sizes.max())

    def get_sobol_normal_sample(self, param_size, sample_size):
        """(sample_size, param_size) float64 standard-normal candidates (test_model.py:493-498)."""
        t = _rec.CandidateTables(self.random_seed, sample_size, self.device).table(int(param_size))
        return t.t().to(torch.float64)

    def _get_sample(self, li, cache, group_idx, n):
        key = (group_idx, n)
        if key not in cache:
            lv = self._levels[li]
            D = int(lv.group_end_host[group_idx] - lv.group_start_host[group_idx])
            cache[key] = self.get_sobol_normal_sample(D, n)
        return cache[key]

    def get_sample(self, group_idx, group_sample_size):
        return self._get_sample(0, self.group_samples, group_idx, group_sample_size)

    def h_get_sample(self, group_idx, group_sample_size):
        return self._get_sample(1, self.h_group_samples, group_idx, group_sample_size)

    def hh_get_sample(self, group_idx, group_sample_size):
        return self._get_sample(2, self.hh_group_samples, group_idx, group_sample_size)

    def _pairs(self, rows, blocks):
        return (torch.as_tensor(rows, dtype=torch.int32, device=self.device).reshape(-1).contiguous(),
                torch.as_tensor(blocks, dtype=torch.int32, device=self.device).reshape(-1).contiguous())

    def _std(self, raw: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """softplus(raw) / 6 in the kernels' own arithmetic (rcb_std_transform): encoder, decoder and fit agree bit for bit."""
        from ._lib import check, ptr, stream
        raw = raw.contiguous()
        out = torch.empty_like(raw) if out is None or out.shape != raw.shape else out
        check(self.engine.lib.rcb_std_transform(ptr(raw), ptr(out), raw.numel(), stream()), "rcb_std_transform")
        return out

    def _scales(self, li=0):
        """(posterior, prior) standard deviations of a level as the REC kernels consume them; the buffers are reused
        from round to round."""
        lv = self._levels[li]
        buf = self.__dict__.setdefault("_scale_bufs", {})
        q = buf[(li, "q")] = self._std(lv.log_scale.data, buf.get((li, "q")))
        p = buf[(li, "p")] = self._std(lv.p_log_scale, buf.get((li, "p")))
        return q, p

    def _sample_group(self, li, row_idx, group_idx, n):
        self._ensure_rec(n)
        lv = self._levels[li]
        q_scale, p_scale = self._scales(li)
        pr, pb = self._pairs([row_idx], [group_idx])
        idx, z, logw = _rec.encode(lv, lv.tables_ptr, self._g_dev, q_scale, p_scale, pr, pb, n, lv.max_D,
                                   apply=False, want_logw=True)
        D = int(lv.group_end_host[group_idx] - lv.group_start_host[group_idx])
        return int(idx.item()), z[0, :D].to(torch.float64), logw[0]

    def sample_group(self, row_idx, group_idx, group_sample_size):
        """A*-code one block (test_model.py:501-533): returns (index, z_i, log_w)."""
        return self._sample_group(0, row_idx, group_idx, group_sample_size)

    def h_sample_group(self, row_idx, group_idx, group_sample_size):
        return self._sample_group(1, row_idx, group_idx, group_sample_size)

    def hh_sample_group(self, row_idx, group_idx, group_sample_size):
        return self._sample_group(2, row_idx, group_idx, group_sample_size)

    def _compress_group(self, li, row_idx, group_idx):
        n = int(np.ceil(2 ** self.bit_per_group))
        self._ensure_rec(n)
        lv = self._levels[li]
        q_scale, p_scale = self._scales(li)
        pr, pb = self._pairs([row_idx], [group_idx])
        _rec.encode(lv, lv.tables_ptr, self._g_dev, q_scale, p_scale, pr, pb, n, lv.max_D, apply=True)
        s, e = int(lv.group_start_host[group_idx]), int(lv.group_end_host[group_idx])
        return int(lv.idx[row_idx, group_idx].item()), lv.sample[row_idx, s:e].clone()

    def compress_group(self, row_idx, group_idx):
        return self._compress_group(0, row_idx, group_idx)

    def h_compress_group(self, row_idx, group_idx):
        return self._compress_group(1, row_idx, group_idx)

    def hh_compress_group(self, row_idx, group_idx):
        return self._compress_group(2, row_idx, group_idx)

    def compress_round(self, blocks: Optional[torch.Tensor] = None, level: int = 0, apply: bool = True):
        """Code one block of every row of a level in a single launch.  With blocks=None each
        row codes its largest-KL not-yet-coded block (test_model.py:809-817).  apply=False scores
        and selects without committing anything (timing / what-if)."""
        from ._lib import check, ptr, stream
        n = int(np.ceil(2 ** self.bit_per_group))
        self._ensure_rec(n)
        lv = self._levels[level]
        bufs = self.__dict__.setdefault("_round_bufs", {})
        if level not in bufs:
            i32 = lambda: torch.empty(lv.rows, dtype=torch.int32, device=self.device)
            bufs[level] = (i32(), i32(), i32())                  # picked blocks, pair rows, pair blocks
        picked, pair_rows, pair_blocks = bufs[level]
        if blocks is None:
            kl = self.engine.group_kl(lv)
            blocks = picked
            check(self.engine.lib.rcb_pick_block(ptr(kl), ptr(lv.coded), ptr(blocks), lv.rows, lv.G, stream()),
                  "rcb_pick_block")
        else:
            blocks = blocks.to(device=self.device, dtype=torch.int32).contiguous()
        q_scale, p_scale = self._scales(level)
        # pairs grouped by block: the kernel scores runs of equal blocks against one pass over the candidate table
        check(self.engine.lib.rcb_rec_order(ptr(blocks), ptr(pair_rows), ptr(pair_blocks), lv.rows, lv.G, stream()),
              "rcb_rec_order")
        _rec.encode(lv, lv.tables_ptr, self._g_dev, q_scale, p_scale, pair_rows, pair_blocks, n, lv.max_D, apply=apply)
        return blocks

    def decode_posteriors(self, indices: np.ndarray, level: int = 0) -> torch.Tensor:
        """Receiver side (the reference has none): rebuild every coded value of a level from
        its transmitted (rows, G) index table, the prior and the seed.  (rows, P), group order."""
        n = int(np.ceil(2 ** self.bit_per_group))
        self._ensure_rec(n)
        lv = self._levels[level]
        rows = torch.arange(lv.rows, device=self.device, dtype=torch.int32).repeat_interleave(lv.G).contiguous()
        blocks = torch.arange(lv.G, device=self.device, dtype=torch.int32).repeat(lv.rows).contiguous()
        idx = torch.as_tensor(np.asarray(indices).astype(np.int32), device=self.device).reshape(-1).contiguous()
        out = torch.zeros(lv.rows, lv.P, device=self.device)
        _rec.decode(lv, lv.tables_ptr, self._std(lv.p_log_scale), rows, blocks, idx, n, out, None)
        return out

    # ------------------------------------------------------------------- training --
    def _adam_config(self, optimizer):
        if optimizer is not self._adam_owner or self._levels[0].adam is None:
            self._adam_owner = optimizer
            for lv in self._levels:
                lv.reset_adam()
        g = optimizer.param_groups[0] if optimizer is not None else {}
        b1, b2 = g.get("betas", (0.9, 0.999))
        return dict(lr=float(g.get("lr", 2e-4)), b1=float(b1), b2=float(b2), eps=float(g.get("eps", 1e-8)))

    def fit_step(self, x, y, epoch, adam_cfg, sample_size=5, eps=None, anneal=None):
        """One fused step: forward, loss, backward, (annealing), Adam (test_model.py:622-635).

        With Philox noise the kernel sequence of a step depends on the epoch only through the noise key and the
        Adam bias corrections, so from the second step of a configuration on it is replayed from a captured CUDA
        graph with those scalars in device memory (rcb_step_state); RECOMBINER_GRAPH=0 keeps every step eager."""
        levels, eng = self._levels, self.engine
        do_anneal = bool((epoch % self.kl_adjust_gap == 0) if anneal is None else anneal)
        for lv in levels:
            if lv.adam is None:
                lv.reset_adam()
        if self.use_graph and eps is None and eng.timer is None and not torch.cuda.is_current_stream_capturing():
            ws = self._fit_step_graph(x, y, epoch, adam_cfg, sample_size, do_anneal)
            if ws is not None:
                return ws
        return self._fit_step_eager(x, y, self._noise(epoch, eps), adam_cfg, sample_size, do_anneal)

    def _fit_step_eager(self, x, y, noise, adam_cfg, S, do_anneal):
        levels, eng = self._levels, self.engine
        rows = levels[0].rows
        ws = eng.forward_features(levels, S, noise)
        # d/dy of N * mean_{n,s,pix,c} (y_pred - y)^2
        coef = 2.0 / (S * eng.pix * eng.out)
        eng.mlp(ws, rows, S, x, mode=1, y=y, coef=coef)
        eng.backward_features(ws, rows, S)
        eng.reduce_samples(levels, ws, S, noise, rows)
        for lv in levels:
            if do_anneal:
                eng.group_kl(lv)            # KL of the pre-step posterior ...
            eng.update(lv, ws, S, noise, with_data_grads=True, adam=adam_cfg, rows=rows)
            if do_anneal:                   # ... beta changes only after this step's gradient (test_model.py:629-634)
                eng.anneal(lv, self.beta_step_size, self.kl_upper_buffer, self.kl_lower_buffer,
                           float(self.bit_per_group))
        return ws

    _GRAPH_CACHE = 8

    def _fit_step_graph(self, x, y, epoch, adam_cfg, S, do_anneal):
        levels, eng = self._levels, self.engine
        t = levels[0].adam["t"]
        if any(lv.adam["t"] != t for lv in levels):
            return None
        # everything a captured step holds by value or by pointer
        xt, x_stride = eng.prepare_x(x)
        key = (S, do_anneal, xt.data_ptr(), x_stride, eng.x_generated, y.data_ptr(), tuple(y.shape),
               adam_cfg["b1"], adam_cfg["b2"], adam_cfg["eps"], self.row_offset, eng.map_generation, eng.half_acts,
               self.beta_step_size, self.kl_upper_buffer, self.kl_lower_buffer, float(self.bit_per_group)) + \
            tuple(v for lv in levels for v in (
                lv.loc.data_ptr(), lv.log_scale.data_ptr(), lv.mask.data_ptr(), lv.sample.data_ptr(),
                lv.p_loc.data_ptr(), lv.p_log_scale.data_ptr(), lv.beta.data_ptr(), lv.coded.data_ptr(),
                lv.adam["m1_loc"].data_ptr(), float(lv.beta_scalar), lv.rows))
        noise = self._noise(epoch)
        entry = self._graphs.get(key)
        if entry is None:
            if key not in self._graphs:          # first sight: run eagerly (sizes workspaces, x cache, TMA maps)
                if len(self._graphs) >= self._GRAPH_CACHE:
                    self._graphs.pop(next(iter(self._graphs)))
                self._graphs[key] = None
                return None
            state = torch.zeros(C.sizeof(StepState), dtype=torch.uint8, device=self.device)      # rcb_step_state
            graph = torch.cuda.CUDAGraph()
            eng.step_state = state
            n0 = COUNTERS["launches"]
            try:
                with torch.cuda.graph(graph):
                    ws = self._fit_step_eager(x, y, noise, adam_cfg, S, do_anneal)
            finally:
                eng.step_state = None
            entry = self._graphs[key] = (graph, state, ws, COUNTERS["launches"] - n0)
            COUNTERS["launches"] = n0
        graph, state, ws, n_kernels = entry
        eng.set_step_state(state, noise.seed, noise.step, adam_cfg, t + 1)
        graph.replay()
        COUNTERS["launches"] += n_kernels
        for lv in levels:
            lv.adam["t"] = t + 1
        return ws

    def _host_x_is_shared(self, x):
        """(every row of a host x holds the same Fourier features, they are the canonical features of the grid): checked
        once per tensor.  Shared -> one row is uploaded; canonical -> nothing is uploaded, the MLP kernel regenerates the
        features from the pixel index (FitEngine.fourier_table)."""
        key = (x.data_ptr(), tuple(x.shape), x._version)
        c = self.__dict__.setdefault("_x_shared_cache", {})
        if key not in c:
            c.clear()
            shared = bool(x.shape[0] == 1 or x.stride(0) == 0 or (x == x[:1]).all())
            eng = self.engine
            canonical = bool(shared and eng.gen_x and eng.tc_mlp and eng.n_f in (16, 18) and eng.x_is_canonical(x[0]))
            c[key] = (shared, canonical)
        return c[key]

    def _stage_inputs(self, x, y):
        """Device-side inputs of a train() call.  CUDA tensors are used as they are.  Host tensors are uploaded on a
        copy stream into one of two persistent staging sets (captured fit-step graphs hold their addresses), so the
        upload of call k+1 travels under the kernels of call k when the host buffers are pinned.  Returns
        (x, y, slot); `_release_inputs(slot)` marks the set reusable once the steps that read it are queued."""
        if x.is_cuda and y.is_cuda:
            return x, y.to(torch.float32).contiguous(), None
        st = self.__dict__.get("_staging")
        if st is None:
            st = self._staging = dict(slot=0, bufs=[None, None], stream=torch.cuda.Stream(device=self.device),
                                      up=[torch.cuda.Event(), torch.cuda.Event()],
                                      done=[torch.cuda.Event(), torch.cuda.Event()])
        b = st["slot"]
        st["slot"] ^= 1
        x, y = x.detach(), y.detach()
        shared, canonical = self._host_x_is_shared(x) if not x.is_cuda else (False, False)
        xs = x[:0] if canonical else (x[:1] if shared else x)          # canonical inputs are not uploaded at all
        bufs = st["bufs"][b]
        if bufs is None or bufs[0].shape != xs.shape or bufs[1].shape != y.shape:
            bufs = st["bufs"][b] = (torch.empty(xs.shape, dtype=torch.float32, device=self.device),
                                    torch.empty(y.shape, dtype=torch.float32, device=self.device))
            # Fresh blocks from the caching allocator may be memory that kernels already queued on the compute stream
            # still read (a tensor freed by Python is reusable in stream order on ITS stream only): the copy stream
            # must not write into them before that work is over.
            st["stream"].wait_stream(torch.cuda.current_stream())
        main = torch.cuda.current_stream()
        with torch.cuda.stream(st["stream"]):
            st["stream"].wait_event(st["done"][b])         # the steps that last read this set are over
            bufs[0].copy_(xs, non_blocking=True)
            bufs[1].copy_(y, non_blocking=True)
            st["up"][b].record(st["stream"])
        main.wait_event(st["up"][b])
        xd = None if canonical else (bufs[0].expand(x.shape[0], -1, -1) if shared else bufs[0])
        return xd, bufs[1], b

    def _release_inputs(self, slot):
        if slot is not None:
            self._staging["done"][slot].record(torch.cuda.current_stream())

    def last_loss_terms(self, sample_size=5):
        """Per-(row, sample) summed squared error of the most recent fit step (device tensor, rows * S)."""
        return self.engine.workspace(self._levels[0].rows, sample_size)["sqerr"]

    def train(self, x=True, y=None, n_epochs=0, optimizer=None, verbose=False, sample_size=5, start_epoch=0):
        """Adam loop of test_model.py:621-635 (`predict(random_seed=epoch)` -> loss -> beta update every
        kl_adjust_gap-th epoch -> step), every step one fused kernel sequence.  x, y may live on the host (pinned
        memory makes the upload asynchronous).  `start_epoch` (extension) offsets the epoch counter, so a caller that
        feeds the loop one step per call keeps the reference's noise keys and annealing cadence."""
        if isinstance(x, bool):          # nn.Module.train(mode) / .eval() compatibility
            return super().train(x)
        cfg = self._adam_config(optimizer)
        x, y, slot = self._stage_inputs(x, y)
        it = range(start_epoch, start_epoch + n_epochs)
        if verbose:
            from tqdm import tqdm
            it = tqdm(it)
        for epoch in it:
            self.fit_step(x, y, epoch, cfg, sample_size)
        self._release_inputs(slot)

    def _distortion(self, x, y):
        with torch.no_grad():
            y_pred = self.predict(x.to(self.device)).cpu()
        return metric(y.cpu().numpy(), y_pred.numpy(), self.dataset)

    def _all_kl_bits(self):
        kls = self.update_annealing_factors(False)
        kls = kls if self.patch else (kls,)
        return [k / np.log(2.) for k in kls]

    def _report(self, x, y, header):
        print(header + " Average Distortion %.4f" % np.mean(self._distortion(x, y)), flush=True)
        kl_bits = np.concatenate([k.reshape(-1) for k in self._all_kl_bits()])
        print("Bits per group: ave %.2f" % kl_bits.mean() + " max %.2f" % kl_bits.max(), flush=True)

    def optimize_posteriors(self, x, y, n_epochs, lr, verbose):
        if verbose:
            self._report(x, y, "Initialization:")
            print(' ')
            print("Start to optimize posteriors...", flush=True)
        optimizer = torch.optim.Adam(self.parameters(), lr=lr)
        self.train(x=x, y=y, n_epochs=n_epochs, optimizer=optimizer, verbose=verbose)
        if verbose:
            self._report(x, y, "Optimization Finished.")

    def _compress_level(self, li, x, y, n_epochs_finetune, verbose, lr, fine_tune_gap, largest_kl):
        """Code every block of one level: per round, one block of each row, then a re-fit of
        everything not yet coded with fresh Adam moments (test_model.py:709-728,806-827)."""
        lv = self._levels[li]
        counter = self._PREFIX[li] + "compressed_num"
        if not hasattr(self, counter):
            setattr(self, counter, 0)
        print_step = set(np.round(np.linspace(0, lv.G, 10)).astype(int).tolist())
        it = range(getattr(self, counter), lv.G)
        if verbose:
            from tqdm import tqdm
            it = tqdm(it)
        for _i in it:
            if largest_kl:
                self.compress_round(level=li)
            else:
                self.compress_round(torch.full((lv.rows,), _i, dtype=torch.int32, device=self.device), level=li)
            setattr(self, counter, getattr(self, counter) + 1)
            if getattr(self, counter) % fine_tune_gap == 0:
                optimizer = torch.optim.Adam(self.parameters(), lr=lr)   # fresh moments each round
                self.train(x, y, n_epochs=n_epochs_finetune, optimizer=optimizer, verbose=False)
            if verbose and _i in print_step:
                kl_bits = self._all_kl_bits()[li]
                open_ = ~lv.coded.bool().cpu().numpy()
                if open_.any():
                    print("Compress progress: %d; " % (100 * getattr(self, counter) / lv.G),
                          "Average Distortion %.4f; " % np.mean(self._distortion(x, y)),
                          "KL in uncompressed groups: MAX %.3f" % kl_bits[open_].max(),
                          "AVE %.3f. " % kl_bits[open_].mean(), flush=True)
        if verbose:
            print(' ')

    def compress_posteriors(self, x, y, n_epochs_finetune, h_n_epochs_finetune=None, hh_n_epochs_finetune=None,
                            verbose=False, lr=2e-4, fine_tune_gap=1, compress_from_group_with_largest_kl=True):
        """Progressive coding, coarsest level first: level 3, level 2, then the per-row level
        (test_model.py:687-856).  Returns the distortion of the fully coded posterior."""
        if verbose:
            print("Start to compress posteriors by A* coding...", flush=True)
        if self.patch:
            self._compress_level(2, x, y, hh_n_epochs_finetune, verbose, lr, fine_tune_gap,
                                 compress_from_group_with_largest_kl)
            self._compress_level(1, x, y, h_n_epochs_finetune, verbose, lr, fine_tune_gap,
                                 compress_from_group_with_largest_kl)
        self._compress_level(0, x, y, n_epochs_finetune, verbose, lr, fine_tune_gap, compress_from_group_with_largest_kl)
        distortion = self._distortion(x, y)
        if verbose:
            print("Optimization Finished. Average Distortion %.4f" % np.mean(distortion), flush=True)
        return distortion
