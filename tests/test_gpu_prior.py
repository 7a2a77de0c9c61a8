"""GPU parity of the prior-training path (north_star (a) with weight gradients, (c) EM
statistics): PriorBNNmodel on the kernels vs the reference goldens."""
import contextlib
import io

import numpy as np
import pytest
import torch

from oracle import cases
from oracle import recombiner_oracle as orc

pytestmark = pytest.mark.gpu


def _model(case, precision="fp32"):
    from recombiner_b200.prior_model import PriorBNNmodel
    from tests.helpers import product_mappings
    shape = case["shape"]
    lt, up = product_mappings(case, "cuda")
    m = PriorBNNmodel(in_dim=shape.dims[0], hidden_dims=shape.dims[1:-1], out_dim=shape.dims[-1],
                      train_size=case["rows"], data_dim=shape.data_dim, pixel_sizes=shape.pixel_sizes,
                      upsample_factors=shape.upsample_factors, latent_dim=shape.latent_dim, patch=shape.patch,
                      patch_nums=shape.patch_nums, hierarchical_patch_nums=shape.hier, device="cuda",
                      layer_scales=shape.layer_scales, paddings=shape.paddings, precision=precision)
    W = shape.n_weights
    with torch.no_grad():
        m._loc_all[:, :W] = case["loc"].cuda()
        m._loc_all[:, W:] = case["lpe_loc"].reshape(case["rows"], -1).cuda()
        m._log_scale_all[:, :W] = case["log_scale"].cuda()
        m._log_scale_all[:, W:] = case["lpe_log_scale"].reshape(case["rows"], -1).cuda()
        for k in ("h_loc", "h_log_scale", "hh_loc", "hh_log_scale"):
            if k in case:
                getattr(m, k).copy_(case[k].cuda())
    return m, lt, up


def _close(a, ref, rtol=2e-3):
    np.testing.assert_allclose(a, ref, rtol=rtol, atol=3e-6 * np.abs(ref).max() + 1e-9)


# Stated tolerances per precision.  fp32 = SIMT parity path.  tf32 = the tcgen05 path the bench times (TF32 / fp16 MMA
# operands, fp32 accumulation): every tensor within 3e-2 * max|reference| (precision policy), scalars and
# gradient norms within 1e-2 relative.
TOL = {"fp32": dict(y=(2e-4, 2e-5), scalar=1e-4, norm=1e-3, grad_max=None),
       "tf32": dict(y=None, scalar=1e-2, norm=1e-2, grad_max=3e-2)}


def _check(a, ref, tol, rtol=2e-3, sub_rtol=None):
    a, ref = np.asarray(a), np.asarray(ref)
    if tol["grad_max"] is not None:
        assert np.abs(a - ref).max() <= tol["grad_max"] * np.abs(ref).max() + 1e-12
    elif sub_rtol is not None:
        np.testing.assert_allclose(a, ref, rtol=sub_rtol, atol=2e-5 * np.abs(ref).max())
    else:
        _close(a, ref, rtol)


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
@pytest.mark.parametrize("name,n_data", [("cifar", 3), ("protein", 4), ("patch2d", 1)])
def test_prior_step_matches_reference(golden, name, n_data, precision):
    """One prior-training forward/backward (prior_model.py:129-200,237-250) against the goldens the unmodified reference
    wrote: reconstruction, loss terms, posterior gradients, gradients of A_l and of the upsampler -- on the SIMT parity
    path and on the tcgen05 path that `train` / bench.py run by default."""
    g = golden("prior_" + name)
    tol = TOL[precision]
    case = cases.make_prior_case(name, n_data)
    m, lt, up = _model(case, precision)
    P = case["prior"]
    pri = (P["loc"], P["scale"], P["lpe_loc"], P["lpe_scale"])
    if case["shape"].patch:
        pri += (P["loc"], P["scale"], P["loc"], P["scale"])
    y_hat = m.forward(case["x"].cuda(), lt, up, True, eps=case["eps"]).cpu().numpy()
    if tol["y"] is not None:
        np.testing.assert_allclose(y_hat, g["y_hat"], rtol=tol["y"][0], atol=tol["y"][1])
    else:
        assert np.abs(y_hat - g["y_hat"]).max() <= 3e-2 * np.abs(g["y_hat"]).max()
    mse, kl, grads = m.loss_and_grads(case["x"], case["y"], pri, lt, up, case["kl_beta"], eps=case["eps"])
    assert float(mse) == pytest.approx(float(g["mse"]), rel=tol["scalar"])
    assert float(kl) == pytest.approx(float(g["kl"]), rel=1e-4)
    assert float(m.calculate_kl(*pri)) == pytest.approx(float(g["kl"]), rel=1e-4)
    for k in ("loc", "log_scale", "lpe_loc", "lpe_log_scale", "h_loc", "h_log_scale", "hh_loc", "hh_log_scale"):
        if k in grads:
            _check(grads[k].cpu().numpy(), g["grad_" + k], tol)
    for i in range(4):
        gf = grads[f"A{i}"].flatten().cpu()
        assert float(gf.double().norm()) == pytest.approx(float(g[f"grad_A{i}_norm"]), rel=tol["norm"])
        _check(gf[::997].numpy(), g[f"grad_A{i}_sub"], tol, sub_rtol=1e-2)
    for k in ("conv1", "conv2", "conv3"):
        for leaf in ("weight", "bias"):
            gf = grads[f"{k}.{leaf}"].flatten().cpu()
            assert float(gf.double().norm()) == pytest.approx(float(g[f"grad_{k}.{leaf}_norm"]), rel=tol["norm"]), (k, leaf)
            _check(gf[::97].numpy(), g[f"grad_{k}.{leaf}_sub"], tol, sub_rtol=1e-2)


@pytest.mark.parametrize("name,n_data", [("cifar", 3), ("protein", 4)])
def test_em_prior_update_matches_reference(golden, name, n_data):
    from recombiner_b200.prior_model import em_prior_update
    g = golden("prior_" + name)
    case = cases.make_prior_case(name, n_data)
    m, _, _ = _model(case)
    p_loc, p_scale, lpe_loc, lpe_scale = em_prior_update(m)
    np.testing.assert_allclose(p_loc.cpu().numpy(), g["em_loc"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(p_scale.cpu().numpy(), g["em_scale"], rtol=1e-5)
    mu, sc = orc.em_prior_update(case["lpe_loc"].reshape(case["rows"], -1), case["lpe_log_scale"].reshape(case["rows"], -1))
    np.testing.assert_allclose(lpe_loc.reshape(-1).cpu().numpy(), mu.numpy(), rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(lpe_scale.reshape(-1).cpu().numpy(), sc.numpy(), rtol=1e-5)


def test_prior_train_decreases_loss_and_moves_mappings():
    """A few Adam steps through `train`: ELBO improves, posteriors and mappings move,
    and the initialisation reproduces the reference's seeded draw order."""
    from recombiner_b200.prior_model import PriorBNNmodel
    from tests.helpers import product_mappings
    case = cases.make_prior_case("cifar", 8)
    shape = case["shape"]
    lt, up = product_mappings(case, "cuda")
    m = PriorBNNmodel(in_dim=32, hidden_dims=[32, 32, 32], out_dim=3, train_size=8, data_dim=2, pixel_sizes=[32, 32],
                      upsample_factors=[16, 16], latent_dim=128, patch=False, patch_nums=None,
                      hierarchical_patch_nums=None, random_seed=42, device="cuda")
    torch.manual_seed(42)
    w_std = np.sqrt(6.0 / 32) / 30.0
    ref_loc = torch.rand(8, 3267) * w_std * 2 - w_std
    ref_lpe = torch.randn(8, 2, 2, 128) * 0.1
    assert torch.equal(m.loc.detach().cpu(), ref_loc) and torch.equal(m.lpe_loc.detach().cpu(), ref_lpe)
    P = case["prior"]
    a0 = lt.A[0].detach().clone()
    w0 = up.conv2.weight.detach().clone()
    mse, kl, elbo = m.train(30, 2e-4, case["x"], case["y"], P["loc"], P["scale"], P["lpe_loc"], P["lpe_scale"],
                            None, None, None, None, lt, up, 1e-8, training_mappings=True)
    assert len(elbo) == 30 and np.isfinite(elbo).all()
    assert np.mean(elbo[-5:]) > np.mean(elbo[:5])
    assert not torch.equal(lt.A[0].detach(), a0) and not torch.equal(up.conv2.weight.detach(), w0)
    assert np.isfinite(mse) and np.isfinite(kl) and kl > 0
