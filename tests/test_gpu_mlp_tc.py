"""Tensor-core fused SIREN MLP (rcb_mlp_tc) against the fp32 SIMT kernel (rcb_mlp) on the same
inputs: prediction, squared error, d pe and all weight/bias gradients.  TF32 chain operands,
fp16 weight-gradient operands (10-bit mantissas, round-to-nearest) through sin(30 z): stated
tolerance 3e-2 * max|reference| per tensor, 2e-3 relative on the summed squared error.  In mode 2
`coef` is the dy scale of the tensor-core kernel (ignored by the SIMT kernel)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _args(lib_args, items, S, pix, out, ld_w, mode, t, coef=0.0, n_f=16):
    from recombiner_b200._lib import ptr
    a = lib_args()
    a.wt, a.xt, a.pe = ptr(t["wt"]), ptr(t["xt"]), ptr(t["pe"])
    a.y, a.dy, a.y_pred = ptr(t["y"]), ptr(t["dy"]), ptr(t["y_pred"])
    a.d_pe, a.d_wt, a.sqerr = ptr(t["d_pe"]), ptr(t["d_wt"]), ptr(t["sqerr"])
    a.pe_base = None
    a.x_row_stride, a.pitch_z, a.pitch_y = 0, 0, 0
    a.items, a.S, a.pix, a.n_f, a.out, a.ld_w, a.mode = items, S, pix, n_f, out, ld_w, mode
    a.ph, a.pw = 1, pix
    a.coef, a.w0 = coef, 30.0
    return a


# weight scale: SIREN initialisation sqrt(6/32)/30 = 0.0144 (prior_model.py:101), i.e. 30*W keeps
# unit gain per layer; larger weights make the network chaotic and amplify ANY rounding
@pytest.mark.parametrize("pix,out,rows,S,wscale,n_f", [(1024, 3, 3, 2, 0.015, 16), (96, 3, 5, 1, 0.02, 16), (800, 1, 2, 3, 0.015, 16),
                                                       (128, 3, 1, 1, 0.01, 16),
                                                       # video INR: 18 Fourier features + 16 encodings = 34 inputs (config.py:101-109)
                                                       (6144, 3, 2, 2, 0.015, 18), (200, 3, 3, 1, 0.02, 18)])
def test_mlp_tc_matches_simt(pix, out, rows, S, wscale, n_f):
    from recombiner_b200 import _lib
    from recombiner_b200._lib import MlpArgs, check, stream
    lib = _lib.load()
    g = torch.Generator().manual_seed(pix + out)
    items = rows * S
    n_w = 32 * (n_f + 17) + 2 * 1056 + out * 33
    ld_w = (n_w + 3) // 4 * 4
    base = dict(wt=torch.zeros(items, ld_w), xt=torch.rand(n_f, pix, generator=g) * 2 - 1,
                pe=torch.randn(items, pix, 16, generator=g) * 0.5, y=torch.rand(rows, pix, out, generator=g),
                dy=torch.randn(items, pix, out, generator=g) * 1e-3)
    base["wt"][:, :n_w] = torch.randn(items, n_w, generator=g) * wscale
    res = {}
    for name, fn in (("simt", lib.rcb_mlp), ("tc", lib.rcb_mlp_tc)):
        for mode in (0, 1, 2):
            t = {k: v.cuda().contiguous() for k, v in base.items()}
            t.update(y_pred=torch.zeros(items, pix, out, device="cuda"), d_pe=torch.zeros(items, pix, 16, device="cuda"),
                     d_wt=torch.zeros(items, ld_w, device="cuda"), sqerr=torch.zeros(items, device="cuda"))
            a = _args(MlpArgs, items, S, pix, out, ld_w, mode, t, coef=2.0 / (S * pix * out) if mode != 2 else 256.0, n_f=n_f)
            check(fn(C.byref(a), stream()), name)
            torch.cuda.synchronize()
            res[(name, mode)] = {k: t[k].cpu().numpy() for k in ("y_pred", "d_pe", "d_wt", "sqerr")}

    def close(key, mode, tol=3e-2):
        ref, got = res[("simt", mode)][key], res[("tc", mode)][key]
        err = np.abs(got - ref).max()
        scale = np.abs(ref).max()
        print(f"[mlp_tc pix={pix} out={out} mode={mode}] {key}: max err {err:.3e} (max |ref| {scale:.3e})")
        assert err <= tol * scale + 1e-12, (key, mode, err, scale)

    close("y_pred", 0)
    for mode in (1, 2):
        close("d_pe", mode)
        close("d_wt", mode)
    ref, got = res[("simt", 1)]["sqerr"], res[("tc", 1)]["sqerr"]
    np.testing.assert_allclose(got, ref, rtol=2e-3)


def test_mlp_tc_is_run_to_run_deterministic():
    """Two groups of a CTA add into the same TMEM weight-gradient accumulators; their MMAs take turns in tile
    order, so repeated launches on the same inputs must agree bit for bit (sharded and unsharded compression are
    compared bit for bit in test_gpu_multi.py)."""
    from recombiner_b200 import _lib
    from recombiner_b200._lib import MlpArgs, check, stream
    lib = _lib.load()
    g = torch.Generator().manual_seed(11)
    rows, S, pix, out = 150, 4, 1024, 3
    items = rows * S
    n_w = 3 * 1056 + out * 33
    ld_w = (n_w + 3) // 4 * 4
    base = dict(wt=torch.zeros(items, ld_w), xt=torch.rand(16, pix, generator=g) * 2 - 1,
                pe=torch.randn(items, pix, 16, generator=g) * 0.5, y=torch.rand(rows, pix, out, generator=g),
                dy=torch.zeros(1))
    base["wt"][:, :n_w] = torch.randn(items, n_w, generator=g) * 0.015
    outs = []
    for _ in range(3):
        t = {k: v.cuda().contiguous() for k, v in base.items()}
        t.update(y_pred=torch.zeros(1, device="cuda"), d_pe=torch.zeros(items, pix, 16, device="cuda"),
                 d_wt=torch.zeros(items, ld_w, device="cuda"), sqerr=torch.zeros(items, device="cuda"))
        a = _args(MlpArgs, items, S, pix, out, ld_w, 1, t, coef=2.0 / (S * pix * out))
        check(lib.rcb_mlp_tc(C.byref(a), stream()), "rcb_mlp_tc")
        torch.cuda.synchronize()
        outs.append((t["d_wt"].clone(), t["d_pe"].clone(), t["sqerr"].clone()))
    for o in outs[1:]:
        for a_, b_ in zip(outs[0], o):
            assert torch.equal(a_, b_)


@pytest.mark.parametrize("dataset", ["cifar", "kodak", "audio", "video", "protein"])
def test_generated_fourier_inputs(dataset):
    """The per-axis table rcb_fourier_table writes from coordinate indices reproduces the loaders' X
    (utils.py:265-298, data/image.py:24-27) to 2e-6 on every modality's grid, and the tensor-core MLP fed by the table
    (`x_tab`: inputs regenerated from the pixel index) is bit-identical to the same kernel reading that X as a tensor."""
    from recombiner_b200 import _lib, utils
    from recombiner_b200._lib import MlpArgs, check, ptr, stream
    from recombiner_b200.config import configs
    from recombiner_b200.engine import FitEngine
    cfg = configs[dataset]
    dims = [cfg["input_dim"]] + cfg["hidden_dims"] + [cfg["output_dim"]]
    eng = FitEngine(dims, cfg["data_dim"], cfg["pixel_sizes"], cfg["upsample_factors"], cfg["latent_dim"],
                    cfg["layerwise_scale_factors"], cfg["paddings"], 30.0, "cuda", patch_nums=cfg["patch_nums"] if cfg["patch"] else None)
    coords, _ = utils.to_grid_coordinates_and_features(torch.zeros(1, *cfg["pixel_sizes"]))
    x1 = utils.fourier_features(coords, cfg["fourier_dim"])                          # the reference's recipe, on the host
    ft = eng.fourier_table()
    assert ft is not None
    err = float((ft["canon"].cpu() - x1).abs().max())
    assert err <= 2e-6, err
    assert eng.x_is_canonical(x1) and not eng.x_is_canonical(x1 * 1.001)
    xt, stride = eng.prepare_x(x1[None].cuda().expand(3, -1, -1))
    assert eng.x_generated and stride == 0
    # kernel level: table lookup == tensor read of the same values
    lib = _lib.load()
    g = torch.Generator().manual_seed(3)
    pix, out, n_f = eng.pix, eng.out, eng.n_f
    items, S = 2, 1
    n_w = eng.W
    ld_w = eng.ldw
    wt = torch.zeros(items, ld_w)
    wt[:, :n_w] = torch.randn(items, n_w, generator=g) * 0.015
    base = dict(wt=wt.cuda(), pe=(torch.randn(items, pix, 16, generator=g) * 0.5).cuda(), y=torch.rand(items, pix, out, generator=g).cuda())
    res = []
    for use_tab in (False, True):
        t = dict(y_pred=torch.zeros(1, device="cuda"), d_pe=torch.zeros(items, pix, 16, device="cuda"),
                 d_wt=torch.zeros(items, ld_w, device="cuda"), sqerr=torch.zeros(items, device="cuda"))
        a = MlpArgs()
        a.wt, a.xt, a.pe, a.y = ptr(base["wt"]), ptr(ft["canon_t"]), ptr(base["pe"]), ptr(base["y"])
        a.d_pe, a.d_wt, a.sqerr = ptr(t["d_pe"]), ptr(t["d_wt"]), ptr(t["sqerr"])
        a.items, a.S, a.pix, a.n_f, a.out, a.ld_w, a.mode = items, S, pix, n_f, out, ld_w, 1
        a.ph, a.pw, a.coef, a.w0 = 1, pix, 2.0 / (pix * out), 30.0
        if use_tab:
            a.x_tab, a.x_axes, a.x_nfreq = ptr(ft["tab"]), eng.data_dim, ft["nf"]
            for i in range(eng.data_dim):
                a.x_size[i], a.x_off[i] = eng.pixel_sizes[i], ft["offs"][i]
        check(lib.rcb_mlp_tc(C.byref(a), stream()), "rcb_mlp_tc")
        torch.cuda.synchronize()
        res.append({k: v.clone() for k, v in t.items()})
    for k in ("d_pe", "d_wt", "sqerr"):
        assert torch.equal(res[0][k], res[1][k]), k
