"""The reference's functional glue under its own names (utils.map_lpe_to_inr_inputs, utils.py:4-120;
utils.map_hierarchical_model_to_int_weights, utils.py:122-198), routed to the kernels, against the oracle."""
import numpy as np
import pytest
import torch

from oracle import cases
from oracle import recombiner_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,n_data", [("cifar", 3), ("patch2d", 2), ("patch1d", 2)])
def test_map_lpe_to_inr_inputs_matches_oracle(name, n_data):
    import utils                                    # the drop-in module name
    from tests.helpers import product_mappings
    case = cases.make_fit_case(name, n_data, 2)
    shape = case["shape"]
    _, up = product_mappings(case, "cuda")
    S, rows = 2, case["rows"]
    g = torch.Generator().manual_seed(5)
    lpe = torch.randn(S, rows, shape.n_latent, generator=g) * 0.3
    ref = orc.latent_to_pe(case["w_up"], lpe, shape).numpy()
    got = utils.map_lpe_to_inr_inputs(up, lpe.cuda(), shape.latent_dim, shape.pixel_sizes, shape.upsample_factors,
                                      shape.patch, shape.patch_nums, shape.data_dim).cpu().numpy()
    assert got.shape == ref.shape == (rows, S, shape.n_pixels, 16)
    # default precision is the tcgen05 path: 3e-2 * max|reference| per tensor (precision policy)
    assert np.abs(got - ref).max() <= 3e-2 * np.abs(ref).max()


def test_map_hierarchical_model_to_int_weights():
    import utils
    shape = cases.shape_of("patch2d")
    rows, W = 2 * 8, shape.n_weights
    g = torch.Generator().manual_seed(9)
    loc, h_loc, hh_loc = torch.randn(rows, W, generator=g), torch.randn(rows // 2, W, generator=g), torch.randn(2, W, generator=g)
    tiny = lambda t: torch.full_like(t, 1e-9)
    args = (shape.hier, shape.patch_nums, shape.data_dim)
    # negligible scales: the sum of the three levels' means with the reference's row expansion (utils.py:151-189)
    hw = utils.map_hierarchical_model_to_int_weights(True, loc.cuda(), tiny(loc).cuda(), h_loc.cuda(), tiny(h_loc).cuda(),
                                                     hh_loc.cuda(), tiny(hh_loc).cuda(), 3, *args).cpu()
    ref = loc + orc.expand_level2(h_loc, shape) + orc.expand_level3(hh_loc, shape)
    assert hw.shape == (rows, 3, W)
    np.testing.assert_allclose(hw.numpy(), ref[:, None].expand(-1, 3, -1).numpy(), rtol=0, atol=2e-6)
    # real scales: per-element mean and variance over many samples match mu and s1^2 + s2^2 + s3^2
    s1, s2, s3 = 0.3, 0.4, 0.5
    S = 4000
    hw = utils.map_hierarchical_model_to_int_weights(True, loc[:, :64].cuda(), torch.full((rows, 64), s1).cuda(),
                                                     h_loc[:, :64].cuda(), torch.full((rows // 2, 64), s2).cuda(),
                                                     hh_loc[:, :64].cuda(), torch.full((2, 64), s3).cuda(), S, *args).cpu()
    mean, var = hw.mean(1), hw.var(1)
    np.testing.assert_allclose(mean.numpy(), ref[:, :64].numpy(), atol=5 * np.sqrt(0.5 / S))
    np.testing.assert_allclose(var.numpy(), np.full((rows, 64), s1 * s1 + s2 * s2 + s3 * s3), rtol=0.15)
    # independent noise per patch at every level (utils.py:179-189): two patches of one level-3 row do not share it
    d = (hw[0] - ref[0, :64]) , (hw[1] - ref[1, :64])
    corr = float((d[0] * d[1]).mean() / (d[0].std() * d[1].std()))
    assert abs(corr) < 0.05
    # single-level form
    hw1 = utils.map_hierarchical_model_to_int_weights(False, loc.cuda(), tiny(loc).cuda(), None, None, None, None, 2, None, None, 2).cpu()
    np.testing.assert_allclose(hw1[:, 0].numpy(), loc.numpy(), atol=2e-6)


@pytest.mark.parametrize("name", ["cifar", "protein"])
def test_upsample_module_forward_matches_oracle(name):
    """`upsample_net(latents)` -- the call the reference makes at utils.py:54,98 -- works on the drop-in module."""
    from tests.helpers import product_mappings
    case = cases.make_fit_case(name, 2, 1)
    shape = case["shape"]
    _, up = product_mappings(case, "cuda")
    g = torch.Generator().manual_seed(2)
    x = torch.randn(5, shape.latent_dim, *shape.lpe_dims, generator=g) * 0.3
    ref = orc.upsample_forward(case["w_up"], x, shape).numpy()
    got = up(x.cuda()).cpu().numpy()
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 3e-2 * np.abs(ref).max()
