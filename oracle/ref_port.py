"""CPU port of the reference's compression loop -- TEST INFRASTRUCTURE / CPU BASELINE.

A class-level restatement (torch CPU autograd + torch.optim.Adam + numpy) of
`TestBNNmodel.train` / `update_annealing_factors` / `compress_posteriors`
(test_model.py:379-439,621-635,800-827) built from the functions of
recombiner_oracle.py.  It keeps the reference's cost structure (global reseed per
step, beta gathered on the host, per-row np.bincount, serial row loop for REC, f64
candidate tables cached per block) so that timing it on the host cores is a fair
stand-in for the reference's CPU path.  Used by bench.py (`cpu_baseline`,
`--impl reference`) and by trajectory tests; never by the product.
"""
from __future__ import annotations

import numpy as np
import torch

from . import recombiner_oracle as orc


class OracleCompressor:
    def __init__(self, case: dict, seed: int = 42, lr: float = 2e-4, kl_adjust_gap: int = 10, bits: float = 16.0):
        self.shape = case["shape"]
        if self.shape.patch:
            raise NotImplementedError("port covers the non-patch modalities used by the bench")
        L = case["lvl1"]
        self.lv = orc.Level(loc=L["loc"].clone().requires_grad_(True), log_scale=L["log_scale"].clone().requires_grad_(True),
                            p_loc=L["p_loc"], p_log_scale=L["p_log_scale"], group_to_param=L["group_to_param"],
                            group_idx=L["group_idx"], group_start=L["group_start"], group_end=L["group_end"],
                            mask=L["mask"].clone(), sample=L["sample"].clone())
        self.beta = L["beta"].clone()
        self.coded = L["coded"].copy()
        self.idx = np.zeros(self.coded.shape)
        self.A, self.w_up = case["A"], case["w_up"]
        self.x, self.y = case["x"], case["y"]
        self.seed, self.lr, self.gap = seed, lr, kl_adjust_gap
        # bits per block (test_model.py:98): candidates per block = ceil(2 ** bits), also the annealing target
        self.bits, self.n_cand = float(bits), int(np.ceil(2 ** bits))
        self.rows = self.lv.loc.shape[0]
        self.W, self.Ln = self.shape.n_weights, self.shape.n_latent
        self._tables = {}
        self._gumbel = None
        self.new_optimizer()

    def new_optimizer(self):
        self.opt = torch.optim.Adam([self.lv.loc, self.lv.log_scale], lr=self.lr)

    def draw_eps(self, epoch: int, S: int):
        torch.manual_seed(epoch)                               # test_model.py:284-285
        lpe = torch.randn(S, self.rows, self.Ln)               # draw order: lpe, then weights
        w = torch.randn(self.rows, S, self.W)
        return {"lpe": lpe, "w": w}

    def fit_step(self, epoch: int, S: int = 5, eps=None):
        eps = eps if eps is not None else self.draw_eps(epoch, S)
        y_pred = orc.predict(self.x, self.lv, self.A, self.w_up, self.shape, eps, S)
        loss = orc.fit_loss(y_pred, self.y) + orc.weighted_kl(self.lv, self.beta)
        if epoch % self.gap == 0:
            self.beta = orc.anneal_beta(self.beta, orc.group_kl_nats(self.lv), self.coded, bits=self.bits)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return float(loss.detach())

    # ---- REC ---------------------------------------------------------------
    def table(self, block: int):
        D = int(self.lv.group_end[block] - self.lv.group_start[block])
        if block not in self._tables:                          # cached per block like test_model.py:459-471
            self._tables[block] = orc.candidate_table(D, self.n_cand, self.seed)
        return self._tables[block]

    def gumbel(self):
        if self._gumbel is None:
            self._gumbel = orc.gumbel_sequence(self.seed, self.n_cand)
        return self._gumbel

    def code_block(self, row: int, block: int):
        s, e = int(self.lv.group_start[block]), int(self.lv.group_end[block])
        with torch.no_grad():
            q_scale = orc.std_transform(self.lv.log_scale[row, s:e]).numpy()
            p_scale = orc.std_transform(self.lv.p_log_scale[s:e]).numpy()
            i, z, _ = orc.rec_encode(self.lv.loc[row, s:e].detach().numpy(), q_scale, self.lv.p_loc[s:e].numpy(),
                                     p_scale, self.table(block), self.gumbel())
        self.idx[row, block] = i
        self.coded[row, block] = True
        self.lv.sample[row, s:e] = torch.from_numpy(z)
        self.lv.mask[row, s:e] = 1.0
        self.beta[row, block] = 0.0
        return i, z

    def compress_round(self):
        """Serial row loop with a full KL recompute per row (test_model.py:807-818)."""
        chosen = []
        for row in range(self.rows):
            bits = orc.group_kl_nats(self.lv)[row] / orc.LN2
            bits[self.coded[row]] = -1e10
            b = int(bits.argmax())
            self.code_block(row, b)
            chosen.append(b)
        return chosen

    def compress(self, n_fit: int, n_finetune: int, S: int = 5):
        """optimize_posteriors + compress_posteriors with explicit (short) schedule."""
        for ep in range(n_fit):
            self.fit_step(ep, S)
        for _ in range(self.lv.n_groups):
            self.compress_round()
            self.new_optimizer()
            for ep in range(n_finetune):
                self.fit_step(ep, S)

    def reconstruct(self, S: int = 1):
        with torch.no_grad():
            eps = {"lpe": torch.randn(S, self.rows, self.Ln), "w": torch.randn(self.rows, S, self.W)}
            return orc.predict(self.x, self.lv, self.A, self.w_up, self.shape, eps, S)
