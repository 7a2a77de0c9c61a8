"""The reference's two command lines, end to end on a temp directory (main_prior_training.py:25-341,
main_compression.py:25-178): `main([...])` of both drivers with the schedule shortened through the documented
environment overrides -- dataset directory in, prior checkpoint (8-object pickle stream) out, CSVs out."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cifar_dir(path, n, seed):
    rs = np.random.RandomState(seed)
    os.makedirs(path, exist_ok=True)
    yy, xx = np.meshgrid(np.linspace(0, 1, 32), np.linspace(0, 1, 32), indexing="ij")
    for i in range(n):
        img = np.stack([0.5 + 0.4 * np.sin(6.28 * (rs.rand() * 2 * yy + rs.rand() * 2 * xx) + rs.rand() * 6) for _ in range(3)])
        np.save(os.path.join(path, "img%04d.npy" % i), img.astype(np.float32))


def test_both_command_lines_run_on_a_directory(tmp_path, monkeypatch, capsys):
    import main_compression          # the top-level modules a reference user would run
    import main_prior_training
    train, test, out = str(tmp_path / "train"), str(tmp_path / "test"), str(tmp_path / "out") + "/"
    _cifar_dir(train, 12, 0)
    _cifar_dir(test, 5, 1)
    os.makedirs(out)
    for k, v in (("RECOMBINER_EM_ITERS", "2"), ("RECOMBINER_FIRST_EPOCHS", "12"), ("RECOMBINER_EPOCHS", "6"),
                 ("RECOMBINER_FIT_EPOCHS", "20"), ("RECOMBINER_FINETUNE_EPOCHS", "1")):
        monkeypatch.setenv(k, v)
    main_prior_training.main(["--train_dir", train, "--train_size", "8", "--dataset", "cifar", "--max_bitrate", "0.5",
                              "--saving_dir", out, "--seed", "3"])
    prior = out + "PRIOR_train_size_8_max_bitrate=0.500.pkl"
    assert os.path.exists(prior) and os.path.exists(out + "LOSS_train_size_8_max_bitrate=0.500.pkl")
    objects = main_compression.load_prior(prior)
    assert len(objects) == 8 and objects[1][0].shape == (3779,)
    # keep the number of REC rounds of this smoke run small: regroup the barely trained prior into ~10 blocks
    import pickle
    from recombiner_b200.prior_model import get_grouping_by_kl
    bits = np.random.RandomState(0).gamma(2.0, 1.0, 3779)
    objects[0] = get_grouping_by_kl(bits * (160.0 / bits.sum()))
    with open(prior, "wb") as f:
        for o in objects:
            pickle.dump(o, f)
    main_compression.main(["--test_dir", test, "--test_idx", "0", "--dataset", "cifar", "--prior_path", prior,
                           "--save_dir", out, "--seed", "3"])
    d = np.loadtxt(out + "Distortion_test_id_0.csv", delimiter=",")
    idx = np.loadtxt(out + "GroupIndex_test_id_0.csv", delimiter=",")
    G = objects[0][5]
    assert d.shape == (5,) and np.isfinite(d).all()
    assert idx.shape == (5, G) and (idx >= 0).all() and (idx < 65536).all() and np.array_equal(idx, np.round(idx))
    assert "Expected bpp" in capsys.readouterr().out
