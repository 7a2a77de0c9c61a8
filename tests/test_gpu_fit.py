"""GPU parity of the fit path (north_star (a)): CUDA kernels through the C ABI and
the drop-in TestBNNmodel API, against the reference goldens and the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import cases
from oracle import recombiner_oracle as orc

pytestmark = pytest.mark.gpu
# stated fp32 tolerances: forward 2e-4 rel, gradients 2e-3 rel (atol 1e-6 of the tensor's max)
FWD = dict(rtol=2e-4, atol=2e-5)


def _grad_close(a, ref, rtol=2e-3):
    np.testing.assert_allclose(a, ref, rtol=rtol, atol=2e-6 * np.abs(ref).max() + 1e-9)


def test_library_loads_on_gpu():
    from recombiner_b200 import _lib
    assert _lib.load().rcb_version() >= 100


@pytest.mark.parametrize("M,N,K", [(70, 99, 99), (256, 1056, 1056), (33, 16, 256), (130, 4096, 512), (5, 33, 33)])
def test_gemm_against_fp32_matmul(M, N, K):
    from recombiner_b200 import _lib
    from recombiner_b200._lib import check, ptr, stream
    lib = _lib.load()
    g = torch.Generator().manual_seed(M * 7 + N)
    lda, ldb = (K + 3) // 4 * 4, (N + 3) // 4 * 4
    A = torch.zeros(M, lda); A[:, :K] = torch.randn(M, K, generator=g)
    B = torch.zeros(K, ldb); B[:, :N] = torch.randn(K, N, generator=g)
    bias = torch.randn(8, generator=g)
    Ad, Bd, bd = A.cuda(), B.cuda(), bias.cuda()
    Cd = torch.zeros(M, ldb, device="cuda")
    check(lib.rcb_gemm(ptr(Ad), lda, ptr(Bd), ldb, ptr(Cd), ldb, M, N, K, None, 1, 0, 0, 0, stream()))
    ref = A[:, :K].double() @ B[:, :N].double()
    np.testing.assert_allclose(Cd[:, :N].cpu().numpy(), ref.numpy(), rtol=1e-4, atol=1e-4)
    # bias + leaky-relu epilogue, accumulate, transposed-A
    check(lib.rcb_gemm(ptr(Ad), lda, ptr(Bd), ldb, ptr(Cd), ldb, M, N, K, ptr(bd), 8, 1, 0, 0, stream()))
    ref2 = torch.nn.functional.leaky_relu(ref + bias[torch.arange(N) % 8].double(), 0.01)
    np.testing.assert_allclose(Cd[:, :N].cpu().numpy(), ref2.numpy(), rtol=1e-4, atol=1e-4)
    ldm = (M + 3) // 4 * 4
    At = torch.zeros(K, ldm); At[:, :M] = A[:, :K].t()
    Atd = At.cuda()
    Cd.fill_(1.0)
    check(lib.rcb_gemm(ptr(Atd), ldm, ptr(Bd), ldb, ptr(Cd), ldb, M, N, K, None, 1, 0, 1, 1, stream()))
    np.testing.assert_allclose(Cd[:, :N].cpu().numpy(), (ref + 1.0).numpy(), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("name,n_data,S", [("cifar", 3, 2), ("protein", 4, 3)])
def test_upsampler_matches_oracle(name, n_data, S):
    """fold + (dense | polyphase) kernels == nearest-upsample + conv of the oracle."""
    from tests.helpers import product_test_model
    case = cases.make_fit_case(name, n_data, S)
    m = product_test_model(case, name)
    lv, eng = m._lv, m.engine
    noise = m._noise(None, case["eps"])
    ws = eng.forward_features(lv, S, noise)
    lpe = ws["lpe"].cpu().view(lv.rows, S, -1).permute(1, 0, 2).contiguous()        # (S, N, L)
    pe_ref = orc.latent_to_pe(case["w_up"], lpe, case["shape"])                      # (N, S, pix, 16)
    pe = ws["pe"].cpu().view(lv.rows, S, eng.pix, 16)
    np.testing.assert_allclose(pe.numpy(), pe_ref.numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name,dataset,n_data,S", [("cifar", "cifar", 3, 2), ("protein", "protein", 4, 3)])
def test_predict_loss_and_grads_match_reference(golden, name, dataset, n_data, S):
    from tests.helpers import product_test_model
    g = golden("fit_" + name)
    case = cases.make_fit_case(name, n_data, S)
    m = product_test_model(case, dataset)
    y = case["y"].cuda()
    y_pred = m.predict(case["x"].cuda(), None, S, eps=case["eps"])
    np.testing.assert_allclose(y_pred.detach().cpu().numpy(), g["y_pred"] if S > 1 else g["y_pred"][:, 0], **FWD)
    mse = torch.mean((y_pred - y[:, None]) ** 2) * y.shape[0]
    kl = m.calculate_kl()
    assert mse.item() == pytest.approx(float(g["mse"]), rel=1e-4)
    assert kl.item() == pytest.approx(float(g["kl"]), rel=1e-4)
    (mse + kl).backward()
    _grad_close(m.loc.grad.cpu().numpy(), g["grad_loc"])
    _grad_close(m.log_scale.grad.cpu().numpy(), g["grad_log_scale"])
    # per-block KL and the annealing rule
    kls = m.update_annealing_factors(True)
    np.testing.assert_allclose(kls, g["group_kl"], rtol=2e-5)
    np.testing.assert_array_equal(m.kl_beta.cpu().numpy(), g["beta_after"])        # exact: same f32 factor, same branches


@pytest.mark.parametrize("name,n_data,S", [("cifar", 3, 2), ("protein", 4, 3)])
def test_fused_step_matches_oracle_adam(name, n_data, S):
    """fit_step (fused loss+backward+Adam, beta annealing) == oracle autograd + torch Adam."""
    from tests.helpers import oracle_level, product_test_model
    case = cases.make_fit_case(name, n_data, S)
    m = product_test_model(case, name)
    lv = oracle_level(case["lvl1"])
    opt = torch.optim.Adam([lv.loc, lv.log_scale], lr=2e-4)
    beta = case["lvl1"]["beta"].clone()
    x, y = case["x"].cuda(), case["y"].cuda()
    cfg = dict(lr=2e-4, b1=0.9, b2=0.999, eps=1e-8)
    m._lv.reset_adam()
    for step in range(3):
        m.fit_step(x, y, step, cfg, sample_size=S, eps=case["eps"], anneal=(step == 0))
        y_pred = orc.predict(case["x"], lv, case["A"], case["w_up"], case["shape"], case["eps"], S)
        loss = orc.fit_loss(y_pred, case["y"]) + orc.weighted_kl(lv, beta)
        if step == 0:
            beta = orc.anneal_beta(beta, orc.group_kl_nats(lv), case["lvl1"]["coded"])
        opt.zero_grad(); loss.backward(); opt.step()
    # Adam's first steps are sign-like (|delta| ~ lr): compare the parameter *updates*
    d_ref = (lv.loc.detach() - case["lvl1"]["loc"]).numpy()
    d_gpu = (m.loc.detach().cpu() - case["lvl1"]["loc"]).numpy()
    assert np.abs(d_gpu - d_ref).max() < 0.02 * 6e-4 + 1e-7
    d_ref = (lv.log_scale.detach() - case["lvl1"]["log_scale"]).numpy()
    d_gpu = (m.log_scale.detach().cpu() - case["lvl1"]["log_scale"]).numpy()
    assert np.abs(d_gpu - d_ref).max() < 0.02 * 6e-4 + 1e-7
    np.testing.assert_array_equal(m.kl_beta.cpu().numpy(), beta.numpy())
    # coded entries must not move
    coded = case["lvl1"]["mask"].bool().numpy()
    sq = float(m.engine.workspace(m._lv.rows, S)["sqerr"].sum().item())
    assert np.isfinite(sq)


def test_philox_noise_statistics_and_reproducibility():
    from tests.helpers import product_test_model
    case = cases.make_fit_case("cifar", 8, 4, coded_frac=0.0)
    m = product_test_model(case, "cifar")
    with torch.no_grad():
        m.loc.zero_(); m.log_scale.fill_(3.0)
    sig = float(torch.nn.functional.softplus(torch.tensor(3.0)) / 6)
    ws = m.engine.forward_features(m._lv, 4, m._noise(5))
    hw = ws["hw"][:, :m.engine.W].clone() / sig
    lpe = ws["lpe"].clone() / sig
    for t in (hw, lpe):
        assert abs(float(t.mean())) < 0.01 and abs(float(t.std()) - 1.0) < 0.01
        assert abs(float((t ** 4).mean()) - 3.0) < 0.1
    ws2 = m.engine.forward_features(m._lv, 4, m._noise(5))
    assert torch.equal(ws2["hw"][:, :m.engine.W] / sig, hw)
    ws3 = m.engine.forward_features(m._lv, 4, m._noise(6))
    assert not torch.equal(ws3["hw"][:, :m.engine.W] / sig, hw)


def test_cpu_device_is_refused():
    from recombiner_b200._lib import KernelError
    from tests.helpers import product_test_model
    case = cases.make_fit_case("cifar", 1, 1)
    with pytest.raises(KernelError):
        product_test_model(case, "cifar", device="cpu")


@pytest.mark.parametrize("name,dataset,n_data,S", [("cifar", "cifar", 3, 2), ("protein", "protein", 4, 3)])
def test_tf32_tensor_core_mode_within_stated_tolerance(golden, name, dataset, n_data, S):
    """precision='tf32': reparam + dense conv1 GEMMs on tcgen05 (10-bit-mantissa operands).
    Stated tolerance: forward <= 5e-3 absolute, loss 1e-3 relative, gradients <= 2 % relative
    L2 error per tensor (fp32 mode: 2e-4 / 1e-4 / 2e-3 elementwise)."""
    from tests.helpers import product_test_model
    g = golden("fit_" + name)
    case = cases.make_fit_case(name, n_data, S)
    m = product_test_model(case, dataset, precision="tf32")
    assert m.engine.tc
    y = case["y"].cuda()
    y_pred = m.predict(case["x"].cuda(), None, S, eps=case["eps"])
    err = np.abs(y_pred.detach().cpu().numpy() - g["y_pred"]).max()
    mse = torch.mean((y_pred - y[:, None]) ** 2) * y.shape[0]
    (mse + m.calculate_kl()).backward()
    rel = {}
    for k, t in (("grad_loc", m.loc.grad), ("grad_log_scale", m.log_scale.grad)):
        ref = g[k]
        rel[k] = float(np.linalg.norm(t.cpu().numpy() - ref) / np.linalg.norm(ref))
    print(f"[tf32 {name}] forward max abs err {err:.3e}; loss rel err {abs(mse.item() - float(g['mse'])) / float(g['mse']):.3e}; "
          f"grad rel L2 {rel}")
    assert err < 5e-3
    assert mse.item() == pytest.approx(float(g["mse"]), rel=1e-3)
    assert all(v < 2e-2 for v in rel.values())


@pytest.mark.parametrize("name,dataset,n_data,S", [("patch2d", "kodak", 1, 2), ("patch1d", "audio", 2, 2),
                                                   ("patch3d", "video", 1, 2)])
def test_patch_modalities_match_reference(golden, name, dataset, n_data, S):
    """Patch modalities: stitched latent grid through the upsampler, three-level hierarchical
    weights with per-patch noise, per-column row permutations (utils.py:60-116,142-191;
    test_model.py:182-208,292-330) -- forward, loss, KL, gradients of all six tensors,
    per-block KL and annealing of every level."""
    from tests.helpers import product_test_model
    g = golden("fit_" + name)
    case = cases.make_fit_case(name, n_data, S)
    m = product_test_model(case, dataset)
    np.testing.assert_array_equal(m.permute_patch_x_g2p, g["perm_g2p"])
    np.testing.assert_array_equal(m.h_permute_patch_x_g2p, g["h_perm_g2p"])
    assert m.bpp == pytest.approx(float(g["bpp"]), rel=1e-12)
    y = case["y"].cuda()
    y_pred = m.predict(case["x"].cuda(), None, S, eps=case["eps"])
    np.testing.assert_allclose(y_pred.detach().cpu().numpy(), g["y_pred"], **FWD)
    mse = torch.mean((y_pred - y[:, None]) ** 2) * y.shape[0]
    kl = m.calculate_kl()
    assert mse.item() == pytest.approx(float(g["mse"]), rel=1e-4)
    assert kl.item() == pytest.approx(float(g["kl"]), rel=1e-4)
    (mse + kl).backward()
    for pre in ("", "h_", "hh_"):
        _grad_close(getattr(m, pre + "loc").grad.cpu().numpy(), g[pre + "grad_loc"])
        _grad_close(getattr(m, pre + "log_scale").grad.cpu().numpy(), g[pre + "grad_log_scale"])
    kls = m.update_annealing_factors(True)
    for k, pre in zip(kls, ("", "h_", "hh_")):
        np.testing.assert_allclose(k, g[pre + "group_kl"], rtol=2e-5)
        np.testing.assert_array_equal(getattr(m, pre + "kl_beta").cpu().numpy(), g[pre + "beta_after"])


def test_patch_fused_step_and_progressive_coding():
    """Patch modality through the fused step and the level-3 -> level-2 -> level-1 coding
    order; every level decodes bit-exactly from its index table."""
    from tests.helpers import product_test_model
    case = cases.make_fit_case("patch2d", 1, 2, coded_frac=0.0, total_bits=64.0)
    m = product_test_model(case, "kodak")
    x, y = case["x"].cuda(), case["y"].cuda()
    before = [lv.loc.detach().clone() for lv in m._levels]
    d = m.compress_posteriors(x, y, n_epochs_finetune=2, h_n_epochs_finetune=2, hh_n_epochs_finetune=2,
                              verbose=False, lr=2e-4)
    assert np.isfinite(d)
    for li, lv in enumerate(m._levels):
        assert bool(lv.coded.all()) and not torch.equal(lv.loc.detach(), before[li])
        idx = (m.compressed_idx_groupwise, m.h_compressed_idx_groupwise, m.hh_compressed_idx_groupwise)[li]
        assert idx.shape == (lv.rows, lv.G)
        assert torch.equal(m.decode_posteriors(idx, level=li), lv.sample)


@pytest.mark.parametrize("name,dataset,S", [("kodak", "kodak", 1), ("audio", "audio", 2), ("video", "video", 1)])
def test_full_size_patch_modalities_match_oracle(name, dataset, S):
    """BASELINE configs 3 and 4 at their real shapes (kodak: 96 patches of 64x64 stitched to a
    32x48 latent grid -> polyphase first stage; audio: 60 patches of 800 samples): forward,
    loss and all gradients against the (golden-pinned) oracle on the same inputs and noise."""
    from tests.helpers import oracle_level, product_test_model
    case = cases.make_fit_case(name, 1, S, total_bits=800.0)
    m = product_test_model(case, dataset)
    assert not m.engine.dense1
    lv = {}
    for key in ("lvl1", "lvl2", "lvl3"):
        lv[key] = oracle_level(case[key])
        if key != "lvl3":
            lv[key].perm_g2p = orc.column_row_permutations(*case[key]["loc"].shape)
    ref = orc.predict(case["x"], lv["lvl1"], case["A"], case["w_up"], case["shape"], case["eps"], S, lv["lvl2"], lv["lvl3"])
    loss_ref = orc.fit_loss(ref, case["y"]) + sum(orc.weighted_kl(lv[k], case[k]["beta"]) for k in lv)
    loss_ref.backward()
    y = case["y"].cuda()
    y_pred = m.predict(case["x"].cuda(), None, S, eps=case["eps"])
    yp = y_pred if S > 1 else y_pred[:, None]
    np.testing.assert_allclose(yp.detach().cpu().numpy(), ref.detach().numpy(), **FWD)
    loss = torch.mean((yp - y[:, None]) ** 2) * y.shape[0] + m.calculate_kl()
    assert loss.item() == pytest.approx(loss_ref.item(), rel=1e-4)
    loss.backward()
    for pre, key in (("", "lvl1"), ("h_", "lvl2"), ("hh_", "lvl3")):
        _grad_close(getattr(m, pre + "loc").grad.cpu().numpy(), lv[key].loc.grad.numpy())
        _grad_close(getattr(m, pre + "log_scale").grad.cpu().numpy(), lv[key].log_scale.grad.numpy())


@pytest.mark.parametrize("name,dataset,n_data,S,precision", [("cifar", "cifar", 6, 3, "fp32"),
                                                           ("cifar", "cifar", 6, 3, "tf32"),
                                                           ("patch2d", "kodak", 1, 2, "fp32")])
def test_graph_replayed_steps_equal_eager_steps(name, dataset, n_data, S, precision):
    """train() replays captured fit steps (noise key and Adam bias corrections from rcb_step_state); the
    posteriors, Adam moments and betas must be bit-identical to launching every step eagerly, across
    annealing steps, an optimizer reset and a change of the coded mask."""
    from tests.helpers import product_test_model
    case = cases.make_fit_case(name, n_data, S, coded_frac=0.0)
    x, y = case["x"].cuda(), case["y"].cuda()
    models = []
    for use_graph in (False, True):
        m = product_test_model(case, dataset, precision=precision)
        m.use_graph = use_graph
        m.kl_adjust_gap = 4
        for rnd in range(2):
            opt = torch.optim.Adam(m.parameters(), lr=2e-4 * (rnd + 1))
            m.train(x=x, y=y, n_epochs=11, optimizer=opt, sample_size=S)
            if rnd == 0:                        # code one block of every row, in place, as compress_round does
                for lv in m._levels:
                    lv.mask[:, : lv.P // 3] = 1
        models.append(m)
    eager, graphed = models
    assert any(v is not None for v in graphed._graphs.values()) and not eager._graphs
    for le, lg in zip(eager._levels, graphed._levels):
        assert le.adam["t"] == lg.adam["t"] == 11
        assert torch.equal(le.loc.data, lg.loc.data) and torch.equal(le.log_scale.data, lg.log_scale.data)
        assert torch.equal(le.beta, lg.beta)
        for k in ("m1_loc", "v_loc", "m1_ls", "v_ls"):
            assert torch.equal(le.adam[k], lg.adam[k])


def test_full_size_kodak_tensor_core_path_against_fp32_path():
    """kodak at its real shape on the tcgen05 path -- polyphase first stage with fp16 output, the persistent x2 kernels on
    the stitched 64 x 96 / 128 x 192 grids (one item per tile, ragged tiles), fp16 activations, general sampling kernels --
    against the fp32 SIMT path of the same model on the same noise: forward <= 5e-3 absolute, gradients <= 2 % relative L2
    per tensor (the tolerance stated for the TF32 mode)."""
    from tests.helpers import product_test_model
    S = 1
    case = cases.make_fit_case("kodak", 1, S, total_bits=800.0)
    x, y = case["x"].cuda(), case["y"].cuda()
    out = {}
    for precision in ("fp32", "tf32"):
        m = product_test_model(case, "kodak", precision=precision)
        if precision == "tf32":
            assert m.engine.f2_half and m.engine.half_acts and not m.engine.half_hw
        y_pred = m.predict(x, None, S, eps=case["eps"])
        yp = y_pred if S > 1 else y_pred[:, None]
        loss = torch.mean((yp - y[:, None]) ** 2) * y.shape[0] + m.calculate_kl()
        loss.backward()
        out[precision] = (yp.detach().cpu().numpy(), loss.item(),
                          [getattr(m, pre + k).grad.cpu().numpy() for pre in ("", "h_", "hh_") for k in ("loc", "log_scale")])
    assert np.abs(out["tf32"][0] - out["fp32"][0]).max() < 5e-3
    assert out["tf32"][1] == pytest.approx(out["fp32"][1], rel=1e-3)
    for a, b in zip(out["tf32"][2], out["fp32"][2]):
        assert np.linalg.norm(a - b) / np.linalg.norm(b) < 2e-2


def test_fp16_weight_gradients_of_the_fit_step_match_the_fp32_ones():
    """Fit step on the tensor-core path with the MLP's weight gradients handed to the reparameterisation GEMM as
    scaled fp16 (default) against the same step with fp32 / TF32 operands: d_hw within 2e-3 relative L2 and 1e-2 of
    the largest entry elementwise (both are 10-bit-mantissa products; only the rounding points differ)."""
    from tests.helpers import product_test_model
    case = cases.make_fit_case("cifar", 6, 3, coded_frac=0.0)
    x, y = case["x"].cuda(), case["y"].cuda()
    cfg = dict(lr=2e-4, b1=0.9, b2=0.999, eps=1e-8)
    out = []
    for half in (True, False):
        m = product_test_model(case, "cifar", precision="tf32")
        m.use_graph = False
        m.engine.half_dwt = half
        ws = m.fit_step(x, y, 3, cfg, sample_size=3)
        assert ws["d_wt_is_half"] == half
        out.append(ws["d_hw"][:, :m.engine.W].clone())
    a, b = out
    assert float((a - b).norm() / b.norm()) < 2e-3
    assert float((a - b).abs().max()) < 1e-2 * float(b.abs().max())
