"""Build product-side (recombiner_b200) models from the shared synthetic cases."""
import contextlib
import io

import numpy as np
import torch

from oracle import cases
from oracle import recombiner_oracle as orc


def product_mappings(case, device):
    from recombiner_b200.prior_model import LinearTransform, Upsample
    shape = case["shape"]
    lt = LinearTransform(shape.dims)
    with torch.no_grad():
        for p, a in zip(lt.A, case["A"]):
            p.copy_(a)
    up = Upsample(shape.data_dim, shape.paddings, shape.layer_scales)
    up.load_state_dict({k: v.clone() for k, v in case["w_up"].items()})
    return lt.to(device), up.to(device)


def product_test_model(case, dataset, device="cuda", quiet=True, precision="fp32"):
    from recombiner_b200.test_model import TestBNNmodel
    shape = case["shape"]
    lt, up = product_mappings(case, device)
    kw = {}
    for pre, key in (("", "lvl1"), ("h_", "lvl2"), ("hh_", "lvl3")):
        if key not in case:
            continue
        L = case[key]
        kw.update({pre + "p_loc": L["p_loc"], pre + "p_log_scale": L["p_log_scale"], pre + "init_log_scale": -4.0,
                   pre + "param_to_group": L["param_to_group"], pre + "group_to_param": L["group_to_param"],
                   pre + "n_groups": L["n_groups"], pre + "group_start_index": L["group_start"],
                   pre + "group_end_index": L["group_end"], pre + "group_idx": L["group_idx"]})
    ctx = contextlib.redirect_stdout(io.StringIO()) if quiet else contextlib.nullcontext()
    with ctx:
        m = TestBNNmodel(in_dim=shape.dims[0], hidden_dims=shape.dims[1:-1], out_dim=shape.dims[-1],
                         number_of_datapoints=case["rows"], upsample_factors=shape.upsample_factors,
                         latent_dim=shape.latent_dim, data_dim=shape.data_dim, pixel_sizes=shape.pixel_sizes,
                         patch=shape.patch, patch_nums=shape.patch_nums, hierarchical_patch_nums=shape.hier,
                         dataset=dataset, linear_transform=lt, upsample_net=up, device=device, random_seed=42,
                         layer_scales=shape.layer_scales, paddings=shape.paddings, precision=precision, **kw)
    with torch.no_grad():
        for lv, key in zip(m._levels, ("lvl1", "lvl2", "lvl3")):
            L = case[key]
            lv.loc.copy_(L["loc"])
            lv.log_scale.copy_(L["log_scale"])
            lv.mask.copy_(L["mask"])
            lv.sample.copy_(L["sample"])
            lv.coded.copy_(torch.from_numpy(L["coded"].astype(np.uint8)))
            lv.beta.copy_(L["beta"])
    return m


def oracle_level(L, requires_grad=True):
    return orc.Level(loc=L["loc"].clone().requires_grad_(requires_grad),
                     log_scale=L["log_scale"].clone().requires_grad_(requires_grad),
                     p_loc=L["p_loc"], p_log_scale=L["p_log_scale"], group_to_param=L["group_to_param"],
                     group_idx=L["group_idx"], group_start=L["group_start"], group_end=L["group_end"],
                     mask=L["mask"].clone(), sample=L["sample"].clone())
