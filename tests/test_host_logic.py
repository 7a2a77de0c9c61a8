"""CPU-side checks of the host mirror: the C ABI surface (header <-> ctypes <-> exported
symbols), config / grouping / Gumbel / metrics against reference goldens.  No GPU needed;
nothing here launches a kernel."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "recombiner_b200.h")


def _header_text():
    txt = open(HEADER).read()
    return re.sub(r"/\*.*?\*/", "", txt, flags=re.S)


def test_library_exports_every_declared_symbol():
    from recombiner_b200 import _lib
    declared = set(re.findall(r"\b(rcb_[a-z0-9_]+)\s*\(", _header_text()))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    lib = _lib.load()                      # CDLL load works without a GPU
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.rcb_version() >= 100


def test_ctypes_structs_mirror_the_header():
    from recombiner_b200 import _lib
    txt = _header_text()
    ctype_of = {"int": ctypes.c_int, "float": ctypes.c_float, "double": ctypes.c_double, "int64_t": ctypes.c_int64}
    for m in re.finditer(r"typedef struct \{(.*?)\}\s*(rcb_[a-z_]+);", txt, flags=re.S):
        body, name = m.group(1), m.group(2)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            is_ptr = "*" in decl
            base = decl.replace("const", "").replace("*", " ").split()
            typ, names = base[0], "".join(base[1:]).split(",")
            for n in names:
                n = n.strip()
                ct = ctypes.c_void_p if is_ptr else ctype_of[typ]
                m_arr = re.fullmatch(r"(\w+)\[(\d+)\]", n)          # fixed-size array member
                if m_arr:
                    n, ct = m_arr.group(1), ct * int(m_arr.group(2))
                fields.append((n, ct))
        struct = _lib.STRUCTS[name]
        assert [(n, t) for n, t in struct._fields_] == fields, name


def test_argument_counts_match_header():
    from recombiner_b200 import _lib
    txt = _header_text()
    for m in re.finditer(r"\b(rcb_[a-z0-9_]+)\s*\((.*?)\)\s*;", txt, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        n = 0 if args in ("void", "") else len(args.split(","))
        assert n == len(_lib.SIGNATURES[name]), name


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from recombiner_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.KernelError):
        _lib.load()


def test_cpu_tensors_are_refused():
    from recombiner_b200 import _lib
    with pytest.raises(_lib.KernelError):
        _lib.ptr(torch.zeros(4))


def test_config_matches_reference_snapshot():
    from recombiner_b200.config import configs
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "config_snapshot.json")))
    assert json.loads(json.dumps(configs)) == ref


@pytest.mark.parametrize("P,total", [(3779, 512.0), (4035, 300.0), (501, 90.0)])
def test_grouping_matches_reference(golden, P, total):
    from oracle import cases
    from recombiner_b200.prior_model import get_grouping_by_kl
    g = golden("grouping")
    gi, gs, ge, g2p, p2g, G, gk, _ = get_grouping_by_kl(cases.synthetic_bits(P, total))
    assert G == int(g[f"P{P}_n"])
    for mine, key in ((gi, "group_idx"), (gs, "start"), (ge, "end"), (g2p, "g2p"), (p2g, "p2g")):
        np.testing.assert_array_equal(mine, g[f"P{P}_{key}"])
    np.testing.assert_allclose(gk, g[f"P{P}_kls"], rtol=1e-12)


def test_gumbel_sequence_matches_reference(golden):
    from recombiner_b200.rec import gumbel_sequence
    g = golden("rec")
    seq = gumbel_sequence(42, 65536)
    np.testing.assert_array_equal(seq[:256], g["gumbel_head"])
    np.testing.assert_array_equal(seq[-256:], g["gumbel_tail"])


def test_metrics_and_inputs_match_reference(golden):
    from recombiner_b200 import utils
    g = golden("misc")
    rs = np.random.RandomState(3)
    a, b = rs.rand(4, 1024, 3), rs.rand(4, 1024, 3) * 1.2 - 0.1
    assert utils.PSNR(a, b, True) == pytest.approx(float(g["psnr_round"]), rel=1e-12)
    assert utils.PSNR(a, b, False) == pytest.approx(float(g["psnr_noround"]), rel=1e-12)
    np.testing.assert_allclose(utils.batch_PSNR(a, b, True), g["batch_psnr"], rtol=1e-12)
    np.testing.assert_allclose(utils.batch_RMSD(a, b, 25), g["batch_rmsd"], rtol=1e-12)
    np.testing.assert_allclose(utils.metric(a, b, "cifar"), g["batch_psnr"], rtol=1e-12)
    for name, sizes, fd in (("cifar", [32, 32], 16), ("protein", [96], 16), ("video", [24, 16, 16], 18)):
        coords, feats = utils.to_grid_coordinates_and_features(torch.zeros(1, *sizes))
        np.testing.assert_allclose(utils.fourier_features(coords, fd).numpy(), g["fourier_" + name], rtol=0, atol=2e-6)
    counts, cum = utils.count_net_params(32, [32, 32, 32], 3)
    assert counts == [1056, 1056, 1056, 99] and int(cum[-1]) == 3267


def test_checkpoint_classes_have_reference_layout():
    from recombiner_b200.prior_model import LinearTransform, Upsample
    lt = LinearTransform([32, 32, 32, 32, 3])
    assert [tuple(a.shape) for a in lt.A] == [(1056, 1056)] * 3 + [(99, 99)]
    assert all(float(a.detach().abs().max()) <= (1.0 + 1e-6) / a.shape[0] for a in lt.A)    # fp32 rounding of 1/n
    up = Upsample(2, [2, 1, 1], [4, 2, 2])
    assert sorted(up.state_dict()) == sorted(f"conv{i}.{k}" for i in (1, 2, 3) for k in ("weight", "bias"))
    assert tuple(up.conv1.weight.shape) == (64, 128, 5, 5) and tuple(up.conv3.weight.shape) == (16, 64, 3, 3)
    assert sum(p.numel() for p in up.parameters()) == 251024
    assert sum(p.numel() for p in Upsample(1, [2, 1, 1], [4, 2, 2]).parameters()) == 56464


def test_bitstream_pack_roundtrip():
    from recombiner_b200 import decode
    rs = np.random.RandomState(1)
    tabs = [rs.randint(0, 65536, (8, 33)), rs.randint(0, 65536, (2, 5)), rs.randint(0, 65536, (1, 7))]
    blob = decode.pack_bitstream(tabs)
    assert len(blob) == 12 + 3 * 8 + 2 * sum(t.size for t in tabs)
    rows, back = decode.unpack_bitstream(blob)
    assert rows == 8 and all(np.array_equal(a, b) for a, b in zip(tabs, back))
    with pytest.raises(ValueError):
        decode.pack_bitstream([np.array([[70000]])])
    with pytest.raises(ValueError):
        decode.unpack_bitstream(b"nope" + blob[4:])


def test_batched_gemm_argument_block():
    """FitEngine._batch_args lays out rcb_gemm_tc_batch's host arrays (pointers, row strides, N, K) in call order."""
    import ctypes as C
    from recombiner_b200.engine import FitEngine
    Bt = [torch.zeros(3, 8), torch.zeros(5, 16)]
    args, keep = FitEngine._batch_args([1024, 2048], 40, Bt, [4096, 8192], 48, 7, [3, 5], [8, 16], 1, 0.5)
    nb, A, lda, B, ldb, Cp, ldc, M, N, K, in_half, out_scale = args
    assert (nb, lda, ldc, M, in_half, out_scale) == (2, 40, 48, 7, 1, 0.5)
    assert list((C.c_void_p * 2).from_address(A)) == [1024, 2048]
    assert list((C.c_void_p * 2).from_address(B)) == [t.data_ptr() for t in Bt]
    assert list((C.c_int * 2).from_address(ldb)) == [8, 16]
    assert list((C.c_void_p * 2).from_address(Cp)) == [4096, 8192]
    assert list((C.c_int * 2).from_address(N)) == [3, 5] and list((C.c_int * 2).from_address(K)) == [8, 16]
