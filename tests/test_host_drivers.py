"""CPU-side checks of the driver-level host logic: KL budget window and global beta controller
(main_prior_training.py:75-83,135-154), the input producers of data/load_data.py against the reference's own
per-modality builders, row sharding of the CLIs, and the unmodified reference (oracle/_ref) against the oracle
restatement on BASELINE config 1 (one cifar-shape image, --device cpu).  No GPU needed."""
import math
import os
import pickle
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_ref():
    from oracle import build_ref
    return build_ref.available()


needs_ref = pytest.mark.skipif(not _have_ref(), reason="oracle/_ref (copy of the reference) not built")


# ------------------------------------------------------------------ E2: budgets and the beta controller --
@pytest.mark.parametrize("dataset,max_bitrate,lo,hi", [
    # hand-evaluated from main_prior_training.py:75-83 with config.py's bitrate_range / lowest_bitrate
    ("cifar", 0.5, 0.2 * 1024, 0.5 * 1024),            # max(0.1, 0.5 - 0.3) * 32*32
    ("cifar", 0.3, 0.1 * 1024, 0.3 * 1024),            # lowest_bitrate wins
    ("kodak", 0.2, 0.1 * 4096, 0.2 * 4096),            # max(0.05, 0.2 - 0.1) * 64*64
    ("audio", 10.0, 9.7 * 800 * (3 / 48000) * 1000, 10.0 * 800 * (3 / 48000) * 1000),    # kbps formula
    ("video", 0.4, 0.1 * 6144, 0.4 * 6144),
    ("protein", 1.0, 0.7 * 96, 1.0 * 96),
])
def test_budgets_follow_the_reference_formula(dataset, max_bitrate, lo, hi):
    from recombiner_b200.config import configs
    from recombiner_b200.main_prior_training import budgets
    got_lo, got_hi = budgets(dataset, configs[dataset], max_bitrate)
    assert got_lo == pytest.approx(lo, rel=1e-12) and got_hi == pytest.approx(hi, rel=1e-12)


def test_step_beta_controller():
    from recombiner_b200.main_prior_training import step_beta
    lo, hi = 100.0, 300.0
    assert step_beta(1e-8, 400.0, lo, hi) == pytest.approx(1.5e-8)          # above the window: x1.5
    assert step_beta(1e-4, 50.0, lo, hi) == pytest.approx(1e-4 / 1.5)       # below: /1.5
    assert step_beta(1e-4, 200.0, lo, hi) == 1e-4                           # inside: unchanged
    assert step_beta(1e-4, 300.0, lo, hi) == 1e-4 and step_beta(1e-4, 100.0, lo, hi) == 1e-4     # edges are inside
    assert step_beta(0.9, 1e9, lo, hi) == 1                                 # clamped above
    assert step_beta(1e-20, 0.0, lo, hi) == 1e-20                           # clamped below
    b = 1e-8                                                                # 550 iterations far above budget saturate at 1
    for _ in range(550):
        b = step_beta(b, 1e9, lo, hi)
    assert b == 1


# ---------------------------------------------------------------- input producers (data/load_data.py) --
def _ref_data_module(name):
    """One of the reference's data/*.py, executed from the packed archive with the reference's utils in scope."""
    import types
    from oracle import build_ref
    ref = build_ref.load()
    saved = sys.modules.get("utils")
    sys.modules["utils"] = ref.utils
    try:
        mod = types.ModuleType("_ref_data_" + name)
        exec(compile(build_ref.read_source("data/%s.py" % name), "oracle/_ref/data/%s.py" % name, "exec"), mod.__dict__)
    finally:
        if saved is not None:
            sys.modules["utils"] = saved
        else:
            del sys.modules["utils"]
    return mod


@needs_ref
def test_input_producers_equal_the_reference_builders(tmp_path):
    from data.load_data import datum_pairs, load_test_set, load_training_set
    g = torch.Generator().manual_seed(0)
    audio = torch.rand(1, 4800, generator=g)
    x, y = datum_pairs(audio, 16, True, [800])
    xr, yr = _ref_data_module("audio").get_audio_pair(audio, 16, True, [800])
    assert torch.equal(x, xr) and torch.equal(y, yr)
    video = torch.rand(24, 3, 32, 32, generator=g)                 # stored (T, C, H, W)
    x, y = datum_pairs(video.permute(1, 0, 2, 3), 18, True, [24, 16, 16])
    xr, yr = _ref_data_module("video").get_video_pair(video, 18, True, [24, 16, 16])
    assert torch.equal(x, xr) and torch.equal(y, yr) and x.shape == (4, 6144, 18)
    prot = torch.rand(3, 96, generator=g)
    x, y = datum_pairs(prot, 16, False, None)
    xr, yr = _ref_data_module("protein").get_protein_pair(prot, 16, False, None)
    assert torch.equal(x[0], xr) and torch.equal(y[0], yr)
    # image directory: portrait file is rotated to landscape, files sorted by name, patches in row-major order
    from PIL import Image
    rs = np.random.RandomState(0)
    d = tmp_path / "imgs"
    d.mkdir()
    for i, shape in enumerate([(128, 64, 3), (64, 128, 3)]):
        Image.fromarray((rs.rand(*shape) * 255).astype(np.uint8)).save(str(d / f"im{i}.png"))
    paths = [str(d / "im0.png"), str(d / "im1.png")]
    x, y = load_test_set(str(d), 1, "kodak", 16, True, [64, 64])
    xr, yr = _ref_data_module("image").load_image(paths[1:2], 16, True, [64, 64])
    assert torch.equal(x, xr) and torch.equal(y, yr) and x.shape == (2, 4096, 16)
    x, y = load_training_set(str(d), "kodak", 3, 10, 16, True, [64, 64])
    idx = np.random.RandomState(3).choice(2, 2, False)
    xr, yr = _ref_data_module("image").load_image([paths[i] for i in idx], 16, True, [64, 64])
    assert torch.equal(x, xr) and torch.equal(y, yr)
    # pickled tensor lists (audio / video / protein): batches of 1000 proteins, single clips otherwise
    with open(tmp_path / "test_dataset.pkl", "wb") as f:
        pickle.dump([torch.rand(3, 96, generator=g) for _ in range(5)], f)
    x, y = load_test_set(str(tmp_path), 0, "protein", 16, False, [96])
    assert x.shape == (5, 96, 16) and y.shape == (5, 96, 3)


def test_input_producers_shapes_without_reference(tmp_path):
    from data.load_data import load_test_set
    np.save(tmp_path / "a.npy", np.random.RandomState(1).rand(3, 32, 32).astype(np.float32))
    np.save(tmp_path / "b.npy", np.random.RandomState(2).rand(32, 32, 3).astype(np.float32))       # HWC is accepted too
    x, y = load_test_set(str(tmp_path), 0, "cifar", 16, False, [32, 32])
    assert x.shape == (2, 1024, 16) and y.shape == (2, 1024, 3)
    assert torch.equal(x[0], x[1])                                   # the Fourier inputs depend on the grid only
    assert float(x.abs().max()) <= 1.0


def test_shard_rows_keeps_whole_data_together():
    from recombiner_b200.parallel import shard_rows
    spans = [shard_rows(96 * 5, 4, r, unit=96) for r in range(4)]
    assert spans == [(0, 192), (192, 288), (288, 384), (384, 480)]      # remainder datum to the first rank
    assert all((b - a) % 96 == 0 for a, b in spans)
    with pytest.raises(ValueError):
        shard_rows(100, 4, 0, unit=96)


# ------------------------------------------- BASELINE config 1: the unmodified reference vs the oracle restatement --
@needs_ref
def test_reference_cpu_arm_agrees_with_the_oracle_on_config1():
    """One synthetic cifar-shape image, small pre-initialised prior, device='cpu' (BASELINE.json configs[0]): the
    UNMODIFIED reference classes, driven through their public API exactly as bench.py's CPU arm drives them, against
    the oracle restatement -- per-block KL, annealed beta, REC index / sample bit-exact."""
    import bench
    from oracle import build_ref, ref_arm
    from oracle import recombiner_oracle as orc
    wl = bench.make_workload(1, seed=5)
    ref = build_ref.load()
    grouping = ref.prior_model.get_grouping_by_kl(wl["bits"])
    mine = orc.grouping_by_kl(wl["bits"])
    for a, b in zip(grouping[:5], mine[:5]):
        np.testing.assert_array_equal(np.asarray(a), np.asarray(b))
    m = ref_arm.build_model(wl["cfg"], "cifar", 1, wl["A"], wl["up"], wl["p_loc"], wl["p_log_scale"], grouping)
    assert m.bpp == pytest.approx(16.0 * grouping[5] / 1024)
    x, y = wl["x"].contiguous(), wl["y"]
    opt = torch.optim.Adam(m.parameters(), lr=2e-4)
    m.train(x=x, y=y, n_epochs=3, optimizer=opt, verbose=False, sample_size=2)       # public loop, global reseed per step
    assert torch.isfinite(m.loc).all() and not torch.equal(m.loc.detach(), wl["p_loc"][None])
    lv = orc.Level(loc=m.loc.detach().clone(), log_scale=m.log_scale.detach().clone(), p_loc=wl["p_loc"],
                   p_log_scale=wl["p_log_scale"], group_to_param=grouping[3], group_idx=grouping[0],
                   group_start=grouping[1], group_end=grouping[2])
    kl_ref = m.update_annealing_factors(False)
    np.testing.assert_allclose(orc.group_kl_nats(lv), kl_ref, rtol=2e-5)
    b = int(np.argmax(kl_ref[0]))
    s, e = int(grouping[1][b]), int(grouping[2][b])
    i_ref, z_ref, logw_ref = m.sample_group(0, b, 65536)
    q_scale = orc.std_transform(lv.log_scale[0, s:e]).numpy()
    p_scale = orc.std_transform(lv.p_log_scale[s:e]).numpy()
    i, z, logw = orc.rec_encode(lv.loc[0, s:e].numpy(), q_scale, wl["p_loc"][s:e].numpy(), p_scale,
                                orc.candidate_table(e - s, 65536, 42), orc.gumbel_sequence(42))
    assert int(i_ref) == int(i)
    np.testing.assert_array_equal(z_ref.numpy().astype(np.float32), z)
    np.testing.assert_allclose(logw, logw_ref.numpy(), rtol=1e-12, atol=1e-9)
