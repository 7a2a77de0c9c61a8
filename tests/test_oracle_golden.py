"""The oracle (oracle/recombiner_oracle.py) against outputs of the unmodified
reference (tests/golden/*.npz, written by oracle/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import cases
from oracle import recombiner_oracle as orc


def _levels(case):
    out = {}
    for key in ("lvl1", "lvl2", "lvl3"):
        if key not in case:
            continue
        L = case[key]
        perm = None
        if case["shape"].patch and key != "lvl3":
            perm = orc.column_row_permutations(*L["loc"].shape)
        out[key] = orc.Level(loc=L["loc"].clone().requires_grad_(True), log_scale=L["log_scale"].clone().requires_grad_(True),
                             p_loc=L["p_loc"], p_log_scale=L["p_log_scale"], group_to_param=L["group_to_param"],
                             group_idx=L["group_idx"], group_start=L["group_start"], group_end=L["group_end"],
                             mask=L["mask"], sample=L["sample"], perm_g2p=perm)
    return out


@pytest.mark.parametrize("name,n_data,S", [("cifar", 3, 2), ("protein", 4, 3), ("patch2d", 1, 2),
                                           ("patch1d", 2, 2), ("patch3d", 1, 2)])
def test_fit_step_matches_reference(golden, name, n_data, S):
    g = golden("fit_" + name)
    case = cases.make_fit_case(name, n_data, S)
    lv = _levels(case)
    y_pred = orc.predict(case["x"], lv["lvl1"], case["A"], case["w_up"], case["shape"], case["eps"], S,
                         lv.get("lvl2"), lv.get("lvl3"))
    np.testing.assert_allclose(y_pred.detach().numpy(), g["y_pred"], rtol=2e-4, atol=2e-5)
    mse = orc.fit_loss(y_pred, case["y"])
    kl = sum(orc.weighted_kl(lv[k], case[k]["beta"]) for k in lv)
    assert mse.item() == pytest.approx(float(g["mse"]), rel=1e-5)
    assert kl.item() == pytest.approx(float(g["kl"]), rel=1e-5)
    (mse + kl).backward()
    for tag, key in (("", "lvl1"), ("h_", "lvl2"), ("hh_", "lvl3")):
        if key not in lv:
            continue
        for nm, t in (("grad_loc", lv[key].loc), ("grad_log_scale", lv[key].log_scale)):
            ref = g[tag + nm]
            np.testing.assert_allclose(t.grad.numpy(), ref, rtol=2e-3, atol=1e-6 * np.abs(ref).max() + 1e-9)
        gk = orc.group_kl_nats(lv[key])
        np.testing.assert_allclose(gk, g[tag + "group_kl"], rtol=1e-6)
        beta = orc.anneal_beta(case[key]["beta"], gk, case[key]["coded"])
        np.testing.assert_array_equal(beta.numpy(), g[tag + "beta_after"])
    if case["shape"].patch:
        np.testing.assert_array_equal(lv["lvl1"].perm_g2p, g["perm_g2p"])
        np.testing.assert_array_equal(lv["lvl2"].perm_g2p, g["h_perm_g2p"])


@pytest.mark.parametrize("name,n_data", [("cifar", 3), ("protein", 4), ("patch2d", 1)])
def test_prior_step_matches_reference(golden, name, n_data):
    g = golden("prior_" + name)
    case = cases.make_prior_case(name, n_data)
    shape = case["shape"]
    leaves = {k: case[k].clone().requires_grad_(True) for k in
              ("loc", "log_scale", "lpe_loc", "lpe_log_scale", "h_loc", "h_log_scale", "hh_loc", "hh_log_scale") if k in case}
    A = [a.clone().requires_grad_(True) for a in case["A"]]
    w_up = {k: v.clone().requires_grad_(True) for k, v in case["w_up"].items()}
    kw = {k: leaves[k] for k in ("h_loc", "h_log_scale", "hh_loc", "hh_log_scale") if k in leaves}
    y_hat = orc.prior_forward(case["x"], leaves["loc"], leaves["log_scale"], leaves["lpe_loc"], leaves["lpe_log_scale"],
                              A, w_up, shape, case["eps"], **kw)
    np.testing.assert_allclose(y_hat.detach().numpy(), g["y_hat"], rtol=2e-4, atol=2e-5)
    mse = torch.mean((y_hat - case["y"]) ** 2) * case["y"].shape[0]
    P = case["prior"]
    kl = orc.gaussian_kl(leaves["loc"], orc.std_transform(leaves["log_scale"]), P["loc"], P["scale"]).sum()
    kl = kl + orc.gaussian_kl(leaves["lpe_loc"], orc.std_transform(leaves["lpe_log_scale"]), P["lpe_loc"], P["lpe_scale"]).sum()
    if shape.patch:
        for t in ("h", "hh"):
            kl = kl + orc.gaussian_kl(leaves[t + "_loc"], orc.std_transform(leaves[t + "_log_scale"]), P["loc"], P["scale"]).sum()
    assert mse.item() == pytest.approx(float(g["mse"]), rel=1e-5)
    assert kl.item() == pytest.approx(float(g["kl"]), rel=1e-5)
    (mse + kl * case["kl_beta"]).backward()
    for k, t in leaves.items():
        ref = g["grad_" + k]
        np.testing.assert_allclose(t.grad.numpy(), ref, rtol=2e-3, atol=1e-6 * np.abs(ref).max() + 1e-9)
    for i, a in enumerate(A):
        gf = a.grad.flatten()
        assert float(gf.double().norm()) == pytest.approx(float(g[f"grad_A{i}_norm"]), rel=1e-4)
        ref = g[f"grad_A{i}_sub"]
        np.testing.assert_allclose(gf[::997].numpy(), ref, rtol=5e-3, atol=1e-5 * np.abs(ref).max())
    for k, p in w_up.items():
        gf = p.grad.flatten()
        assert float(gf.double().norm()) == pytest.approx(float(g["grad_" + k + "_norm"]), rel=1e-4)
    mu, sc = orc.em_prior_update(case["loc"], case["log_scale"])
    np.testing.assert_allclose(mu.numpy(), g["em_loc"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(sc.numpy(), g["em_scale"], rtol=1e-6)


def test_gumbel_sequence(golden):
    g = golden("rec")
    seq = orc.gumbel_sequence(42)
    np.testing.assert_array_equal(seq[:256], g["gumbel_head"])
    np.testing.assert_array_equal(seq[-256:], g["gumbel_tail"])
    assert np.all(np.diff(seq) < 0)


def test_ndtri_restatement_matches_scipy():
    from scipy import special
    u = np.random.RandomState(0).rand(400000)
    u = np.concatenate([u, 10.0 ** -np.arange(1, 12, 0.05), 1 - 10.0 ** -np.arange(1, 12, 0.25)])
    a, b = orc.ndtri_cephes(u), special.ndtri(u)
    assert np.abs((a - b) / b).max() < 2e-15


@pytest.mark.parametrize("D", [1, 3, 25, 114, 136])
def test_candidate_table_and_rec(golden, D):
    g = golden("rec")
    table = orc.candidate_table(D, 65536, 42)
    ref_rows = g[f"table{D}_rows"]
    mism = (table[::257] != ref_rows).sum()
    assert mism == 0, f"{mism} table entries differ from scipy path"
    np.testing.assert_allclose(table.sum(0), g[f"table{D}_colsum"], rtol=0, atol=1e-6)
    rc = cases.make_rec_case(D)
    i, z, lw = orc.rec_encode(rc["q_loc"], rc["q_scale"], rc["p_loc"], rc["p_scale"], table, orc.gumbel_sequence(42))
    assert i == int(g[f"rec{D}_idx"])
    np.testing.assert_array_equal(z, g[f"rec{D}_z"])
    np.testing.assert_allclose(lw[::64], g[f"rec{D}_logw_sub"], rtol=1e-12, atol=1e-9)
    np.testing.assert_array_equal(orc.rec_decode(i, rc["p_loc"], rc["p_scale"], table), z)


@pytest.mark.parametrize("P,total", [(3779, 512.0), (4035, 300.0), (501, 90.0)])
def test_grouping(golden, P, total):
    g = golden("grouping")
    gi, gs, ge, g2p, p2g, G, gk, _ = orc.grouping_by_kl(cases.synthetic_bits(P, total))
    assert G == int(g[f"P{P}_n"])
    for mine, key in ((gi, "group_idx"), (gs, "start"), (ge, "end"), (g2p, "g2p"), (p2g, "p2g")):
        np.testing.assert_array_equal(mine, g[f"P{P}_{key}"])
    np.testing.assert_allclose(gk, g[f"P{P}_kls"], rtol=1e-12)


def test_fourier_inputs(golden):
    g = golden("misc")
    for name, sizes, fd in (("cifar", [32, 32], 16), ("protein", [96], 16), ("video", [24, 16, 16], 18)):
        np.testing.assert_allclose(orc.fourier_inputs(sizes, fd).numpy(), g["fourier_" + name], rtol=0, atol=2e-6)
