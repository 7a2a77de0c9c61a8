"""The oracle port against the UNMODIFIED reference through a whole (short) compression on the CPU.

The golden vectors pin the oracle step by step; this pins the SCHEDULE the port restates -- `optimize_posteriors`
followed by `compress_posteriors` (test_model.py:637-685, 687-856): per round every row codes its largest-KL open block,
Adam is rebuilt, a few fine-tune steps follow with the noise of `predict(random_seed=epoch)` and the beta update of every
kl_adjust_gap-th epoch.  Both sides draw their noise from torch's CPU generator seeded with the epoch (test_model.py:
284-285), so the comparison is not statistical: the chosen blocks, the transmitted indices, the coded samples and the
posterior means must agree exactly after every round.  (This is what the CUDA path's trajectory tests lean on when
they compare against the port, and it is how the port was shown not to be the source of the post-coding PSNR gap reported
in DESIGN.md section 2.)
"""
import os

import numpy as np
import pytest
import torch


def _have_ref():
    from oracle import build_ref
    return build_ref.available()


pytestmark = pytest.mark.skipif(not _have_ref(), reason="oracle/_ref (copy of the reference) not built")


def test_port_equals_reference_through_progressive_coding():
    import bench
    from oracle import recombiner_oracle as orc
    from oracle import ref_arm
    from oracle.ref_port import OracleCompressor
    rows, n_fit, n_finetune, bits = 2, 12, 3, 9.0     # 512 candidates per block instead of 65 536: seconds, not minutes
    saved = bench.TOTAL_BITS
    bench.TOTAL_BITS = 96.0            # six blocks of 16 bits: the whole schedule runs in seconds on the host
    try:
        wl = bench.make_workload(rows, seed=5)
        case = bench.oracle_case(wl)
    finally:
        bench.TOTAL_BITS = saved
    G = case["lvl1"]["n_groups"]
    assert 3 <= G <= 12
    oc = OracleCompressor(case, bits=bits)
    gi, gs, ge, g2p, p2g, _ = orc.grouping_by_kl(wl["bits"])[:6]
    m = ref_arm.build_model(wl["cfg"], "cifar", rows, wl["A"], wl["up"], wl["p_loc"], wl["p_log_scale"], (gi, gs, ge, g2p, p2g, G))
    m.bit_per_group = bits               # attribute of the unmodified class (test_model.py:98): candidates and annealing target
    x, y = wl["x"].contiguous(), wl["y"]
    torch.set_num_threads(min(8, os.cpu_count() or 1))

    def ref_epoch(opt, ep):            # the body of TestBNNmodel.train (test_model.py:622-635) for one epoch
        y_pred = m.predict(x=x, random_seed=ep, sample_size=5)
        elbo = torch.mean((y_pred - y[:, None, :, :]) ** 2) * y.shape[0] + m.calculate_kl()
        if ep % m.kl_adjust_gap == 0:
            m.update_annealing_factors(update=True)
        opt.zero_grad()
        elbo.backward()
        opt.step()
        return float(elbo.detach())

    opt = torch.optim.Adam(m.parameters(), lr=2e-4)
    for ep in range(n_fit):
        lp, lr_ = oc.fit_step(ep, 5), ref_epoch(opt, ep)
        assert lp == pytest.approx(lr_, rel=1e-6)
    assert torch.equal(oc.lv.loc.detach(), m.loc.detach()) and torch.equal(oc.lv.log_scale.detach(), m.log_scale.detach())

    for rnd in range(G):
        chosen = oc.compress_round()
        oc.new_optimizer()
        ref_chosen = []
        for row in range(rows):        # test_model.py:807-818
            kl_bits = m.update_annealing_factors(False)[row] / np.log(2.)
            kl_bits[m.compressed_mask_groupwise[row]] = -1e10
            ref_chosen.append(int(kl_bits.argmax()))
            m.compress_group(row, ref_chosen[-1])
        opt = torch.optim.Adam(m.parameters(), lr=2e-4)
        for ep in range(n_finetune):
            oc.fit_step(ep, 5)
            ref_epoch(opt, ep)
        assert chosen == ref_chosen, (rnd, chosen, ref_chosen)
        np.testing.assert_array_equal(oc.idx, np.asarray(m.compressed_idx_groupwise), err_msg=f"round {rnd}")
        assert torch.equal(oc.lv.sample, m.compressed_sample), rnd
        assert torch.equal(oc.lv.loc.detach(), m.loc.detach()), rnd
        assert torch.equal(oc.lv.log_scale.detach(), m.log_scale.detach()), rnd
    assert bool(np.asarray(m.compressed_mask_groupwise).all()) and bool(oc.coded.all())
