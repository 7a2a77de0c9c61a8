"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

The reference is imported from /root/reference (never copied); inputs come from
oracle/cases.py seeds; noise is injected by temporarily replacing
`torch.randn_like` with a replay of the case's eps tensors in the reference's draw
order (test_model.py:303, utils.py:148/196,181,189; prior_model.py:145).
Only outputs are stored.  TEST INFRASTRUCTURE ONLY.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("RECOMBINER_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
import test_model as ref_test  # noqa: E402  (reference)
import prior_model as ref_prior  # noqa: E402  (reference)
import utils as ref_utils  # noqa: E402  (reference)
import config as ref_config  # noqa: E402  (reference)
sys.path.pop(0)
sys.path.insert(0, ROOT)
from oracle import cases  # noqa: E402
from oracle import recombiner_oracle as orc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(8)


@contextlib.contextmanager
def replay_noise(tensors):
    queue = list(tensors)
    real = torch.randn_like

    def fake(t, **kw):
        e = queue.pop(0)
        assert e.shape == t.shape, (e.shape, t.shape)
        return e.clone()

    torch.randn_like = fake
    try:
        yield
    finally:
        torch.randn_like = real
    assert not queue, "unused noise tensors"


def ref_mappings(case):
    shape = case["shape"]
    lt = ref_prior.LinearTransform(shape.dims)
    for p, a in zip(lt.A, case["A"]):
        p.data.copy_(a)
    up = ref_prior.Upsample(shape.data_dim, shape.paddings, shape.layer_scales)
    sd = {k: v.clone() for k, v in case["w_up"].items()}
    up.load_state_dict(sd)
    return lt, up


def build_ref_test_model(case, dataset):
    shape = case["shape"]
    lt, up = ref_mappings(case)
    kw = {}
    for tag, key in (("", "lvl1"), ("h_", "lvl2"), ("hh_", "lvl3")):
        if key not in case:
            continue
        L = case[key]
        kw.update({tag + "p_loc": L["p_loc"], tag + "p_log_scale": L["p_log_scale"],
                   tag + "init_log_scale": -4.0,
                   tag + "param_to_group": L["param_to_group"], tag + "group_to_param": L["group_to_param"],
                   tag + "n_groups": L["n_groups"], tag + "group_start_index": L["group_start"],
                   tag + "group_end_index": L["group_end"], tag + "group_idx": L["group_idx"]})
    with contextlib.redirect_stdout(io.StringIO()):
        m = ref_test.TestBNNmodel(in_dim=shape.dims[0], hidden_dims=shape.dims[1:-1], out_dim=shape.dims[-1],
                                  number_of_datapoints=case["rows"], upsample_factors=shape.upsample_factors,
                                  latent_dim=shape.latent_dim, data_dim=shape.data_dim, pixel_sizes=shape.pixel_sizes,
                                  patch=shape.patch, patch_nums=shape.patch_nums, hierarchical_patch_nums=shape.hier,
                                  dataset=dataset, linear_transform=lt, upsample_net=up, device="cpu",
                                  random_seed=42, **kw)
    for tag, key in (("", "lvl1"), ("h_", "lvl2"), ("hh_", "lvl3")):
        if key not in case:
            continue
        L = case[key]
        getattr(m, tag + "loc").data.copy_(L["loc"])
        getattr(m, tag + "log_scale").data.copy_(L["log_scale"])
        setattr(m, tag + "compressed_mask", L["mask"].clone())
        setattr(m, tag + "compressed_sample", L["sample"].clone())
        setattr(m, tag + "compressed_mask_groupwise", L["coded"].copy())
        setattr(m, tag + "kl_beta", L["beta"].clone())
    return m


def fit_golden(name, dataset, n_data, S, **kw):
    case = cases.make_fit_case(name, n_data, S, **kw)
    m = build_ref_test_model(case, dataset)
    eps = case["eps"]
    order = [eps["lpe"], eps["w"]] + ([eps["h"], eps["hh"]] if case["shape"].patch else [])
    with replay_noise(order):
        y_pred = m.predict(case["x"], None, S)
    if S == 1:
        y_pred = y_pred[:, None]
    mse = torch.mean((y_pred - case["y"][:, None]) ** 2) * case["y"].shape[0]
    kl = m.calculate_kl()
    (mse + kl).backward()
    out = dict(y_pred=y_pred.detach().numpy(), mse=mse.item(), kl=kl.item())
    for tag in ("", "h_", "hh_"):
        if hasattr(m, tag + "loc") and isinstance(getattr(m, tag + "loc"), torch.nn.Parameter):
            out[tag + "grad_loc"] = getattr(m, tag + "loc").grad.numpy()
            out[tag + "grad_log_scale"] = getattr(m, tag + "log_scale").grad.numpy()
    kls = m.update_annealing_factors(True)
    if case["shape"].patch:
        out["group_kl"], out["h_group_kl"], out["hh_group_kl"] = kls
        out["h_beta_after"] = m.h_kl_beta.numpy()
        out["hh_beta_after"] = m.hh_kl_beta.numpy()
        out["perm_g2p"] = m.permute_patch_x_g2p.astype(np.int32)
        out["h_perm_g2p"] = m.h_permute_patch_x_g2p.astype(np.int32)
    else:
        out["group_kl"] = kls
    out["beta_after"] = m.kl_beta.numpy()
    out["bpp"] = float(m.bpp)
    np.savez_compressed(os.path.join(OUT, f"fit_{name}.npz"), **out)
    print("fit", name, "rows", case["rows"], "mse", out["mse"], "kl", out["kl"])


def prior_golden(name, n_data):
    case = cases.make_prior_case(name, n_data)
    shape = case["shape"]
    lt, up = ref_mappings(case)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ref_prior.PriorBNNmodel(in_dim=shape.dims[0], hidden_dims=shape.dims[1:-1], out_dim=shape.dims[-1],
                                    train_size=case["rows"], data_dim=shape.data_dim, pixel_sizes=shape.pixel_sizes,
                                    upsample_factors=shape.upsample_factors, latent_dim=shape.latent_dim,
                                    patch=shape.patch, patch_nums=shape.patch_nums,
                                    hierarchical_patch_nums=shape.hier, device="cpu")
    for k in ("loc", "log_scale", "lpe_loc", "lpe_log_scale", "h_loc", "h_log_scale", "hh_loc", "hh_log_scale"):
        if k in case:
            getattr(m, k).data.copy_(case[k])
    eps = case["eps"]
    order = [eps["lpe"], eps["w"]] + ([eps["h"], eps["hh"]] if shape.patch else [])
    with replay_noise(order):
        y_hat = m.forward(case["x"], lt, up, True)
    P = case["prior"]
    mse = torch.mean((y_hat - case["y"]) ** 2) * case["y"].shape[0]
    h = (P["loc"], P["scale"]) if shape.patch else (None, None)
    kl = m.calculate_kl(P["loc"], P["scale"], P["lpe_loc"], P["lpe_scale"], h[0], h[1], h[0], h[1])
    (mse + kl * case["kl_beta"]).backward()
    out = dict(y_hat=y_hat.detach().numpy(), mse=mse.item(), kl=kl.item())
    for k in ("loc", "log_scale", "lpe_loc", "lpe_log_scale", "h_loc", "h_log_scale", "hh_loc", "hh_log_scale"):
        if k in case:
            out["grad_" + k] = getattr(m, k).grad.numpy()
    for i, a in enumerate(lt.A):
        gflat = a.grad.flatten()
        out[f"grad_A{i}_sub"] = gflat[::997].numpy()
        out[f"grad_A{i}_norm"] = float(gflat.double().norm())
    for k, p in up.named_parameters():
        gflat = p.grad.flatten()
        out["grad_" + k + "_sub"] = gflat[::97].numpy()
        out["grad_" + k + "_norm"] = float(gflat.double().norm())
    # EM prior update exactly as main_prior_training.py:157-159
    pl = m.loc.clone().detach().mean(0)
    ps = ((m.st(m.log_scale.clone().detach()) ** 2).mean(0) + m.loc.clone().detach().var(0)) ** 0.5
    out["em_loc"], out["em_scale"] = pl.numpy(), ps.numpy()
    np.savez_compressed(os.path.join(OUT, f"prior_{name}.npz"), **out)
    print("prior", name, "mse", out["mse"], "kl", out["kl"])


def rec_golden():
    case = cases.make_fit_case("cifar", 1, 1, coded_frac=0.0)
    m = build_ref_test_model(case, "cifar")
    m.get_gumbel_sample()
    g = m.g_samples.numpy()
    out = dict(gumbel_head=g[:256], gumbel_tail=g[-256:], gumbel_sum=np.array(g.sum()))
    for D in (1, 3, 25, 114, 136):
        t = m.get_sobol_normal_sample(D, 65536).numpy()
        assert t.dtype == np.float64
        # scipy's f32 loop: every entry is an f32 value
        assert np.array_equal(t, t.astype(np.float32).astype(np.float64))
        out[f"table{D}_rows"] = t[::257]                       # 256 rows incl. row 0
        out[f"table{D}_colsum"] = t.sum(0)
        out[f"table{D}_absmax"] = np.array(np.abs(t).max())
        rc = cases.make_rec_case(D)
        # drive the reference's own sample_group on a 1-block model
        mm = build_ref_test_model(case, "cifar")
        mm.g_samples = m.g_samples
        mm.group_start_index = np.array([0]); mm.group_end_index = np.array([D])
        mm.p_loc = torch.from_numpy(rc["p_loc"]); mm.st = lambda v: v          # pass scales directly
        mm.p_log_scale = torch.from_numpy(rc["p_scale"])
        mm.loc = torch.nn.Parameter(torch.from_numpy(rc["q_loc"])[None])
        mm.log_scale = torch.nn.Parameter(torch.from_numpy(rc["q_scale"])[None])
        i, z, lw = mm.sample_group(0, 0, 65536)
        out[f"rec{D}_idx"] = np.array(i)
        out[f"rec{D}_z"] = z.float().numpy()                    # f32 on store (test_model.py:591)
        out[f"rec{D}_z64"] = z.numpy()
        out[f"rec{D}_logw_sub"] = lw.numpy()[::64]
        out[f"rec{D}_logw_max"] = np.array(lw.max().item())
        out[f"rec{D}_logw_top"] = np.sort(lw.numpy())[-8:]
    np.savez_compressed(os.path.join(OUT, "rec.npz"), **out)
    print("rec done; idx", {D: int(out[f'rec{D}_idx']) for D in (1, 3, 25, 114, 136)})


def grouping_golden():
    out = {}
    for P, total in ((3779, 512.0), (4035, 300.0), (501, 90.0)):
        bits = cases.synthetic_bits(P, total)
        gi, gs, ge, g2p, p2g, G, gk, w = ref_prior.get_grouping_by_kl(bits)
        out[f"P{P}_group_idx"] = gi; out[f"P{P}_start"] = gs; out[f"P{P}_end"] = ge
        out[f"P{P}_g2p"] = g2p; out[f"P{P}_p2g"] = p2g; out[f"P{P}_n"] = np.array(G); out[f"P{P}_kls"] = gk
    np.savez_compressed(os.path.join(OUT, "grouping.npz"), **out)
    print("grouping done")


def misc_golden():
    """config snapshot, metrics, Fourier inputs, layer counts."""
    import json
    snap = json.loads(json.dumps(ref_config.configs))        # tuples -> lists
    with open(os.path.join(OUT, "config_snapshot.json"), "w") as f:
        json.dump(snap, f, indent=1, sort_keys=True)
    out = {}
    rs = np.random.RandomState(3)
    a, b = rs.rand(4, 1024, 3), rs.rand(4, 1024, 3) * 1.2 - 0.1
    out["psnr_round"] = np.array(ref_utils.PSNR(a, b, True)); out["psnr_noround"] = np.array(ref_utils.PSNR(a, b, False))
    out["batch_psnr"] = ref_utils.batch_PSNR(a, b, True); out["batch_rmsd"] = ref_utils.batch_RMSD(a, b, 25)
    for name, sizes, fd in (("cifar", [32, 32], 16), ("protein", [96], 16), ("video", [24, 16, 16], 18)):
        datum = torch.zeros(1, *sizes)
        coords, _ = ref_utils.to_grid_coordinates_and_features(datum)
        w = torch.exp(torch.linspace(0, np.log(1024), fd // (2 * len(sizes))))
        inp = torch.matmul(coords.unsqueeze(-1), w.unsqueeze(0)).view(*coords.shape[:-1], -1)
        out["fourier_" + name] = torch.cat([torch.cos(np.pi * inp), torch.sin(np.pi * inp)], -1).numpy()
    np.savez_compressed(os.path.join(OUT, "misc.npz"), **out)
    print("misc done")


if __name__ == "__main__":
    fit_golden("cifar", "cifar", 3, 2)
    fit_golden("protein", "protein", 4, 3)
    fit_golden("patch2d", "kodak", 1, 2)
    fit_golden("patch1d", "audio", 2, 2)
    fit_golden("patch3d", "video", 1, 2)
    prior_golden("cifar", 3)
    prior_golden("protein", 4)
    prior_golden("patch2d", 1)
    rec_golden()
    grouping_golden()
    misc_golden()
