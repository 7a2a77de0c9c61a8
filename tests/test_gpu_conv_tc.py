"""Tensor-core (tcgen05 + 5-D TMA gather) polyphase up-convolutions against the fp32 SIMT
engine on the same inputs and folded weights.  TF32 operands: stated tolerance 2e-3 of
|a|.|b| per output (10-bit mantissas), forward and data gradient; 1-D, 2-D and 3-D grids,
ragged tiles, the 64-byte-swizzle path (oc = 16) and the multi-CTA phase split (16 phases)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GEOMS = {
    # name: (d, h, w, fz, fy, fx, kz, ky, kx, ic, oc, items)
    "cifar_conv2": (1, 8, 8, 1, 2, 2, 1, 3, 3, 64, 64, 7),
    "cifar_conv3": (1, 16, 16, 1, 2, 2, 1, 3, 3, 64, 16, 5),
    "conv1_poly_2d": (1, 6, 10, 1, 4, 4, 1, 5, 5, 128, 64, 3),
    "protein_conv2": (1, 1, 24, 1, 1, 2, 1, 1, 3, 64, 64, 11),
    "audio_conv3_ragged": (1, 1, 300, 1, 1, 2, 1, 1, 3, 64, 16, 2),
    "wide_2d_ragged": (1, 5, 200, 1, 2, 2, 1, 3, 3, 64, 64, 1),
    "video_conv2_3d": (3, 4, 4, 2, 2, 2, 3, 3, 3, 64, 64, 4),
    "video_conv1_3d": (1, 2, 2, 6, 4, 4, 5, 5, 5, 128, 64, 2),
    # persistent x2 kernel: ragged tiles, more than two tiles per CTA (both accumulator sets re-used)
    "f2_many_tiles": (1, 40, 44, 1, 2, 2, 1, 3, 3, 64, 16, 20),
    "f2_oc32": (1, 16, 24, 1, 2, 2, 1, 3, 3, 32, 32, 3),
    # persistent 64 -> 64 fp16 kernel: one item per tile (ragged), two 8 x 8 items per tile (odd count, > 2 tiles per CTA)
    "f2w_ragged": (1, 40, 44, 1, 2, 2, 1, 3, 3, 64, 64, 5),
    "f2w_8x8_many": (1, 8, 8, 1, 2, 2, 1, 3, 3, 64, 64, 701),
}


@pytest.mark.parametrize("name", list(GEOMS))
def test_upconv_tc_matches_simt(name):
    from recombiner_b200 import _lib
    from recombiner_b200._lib import UpconvGeom, check, ptr, stream
    lib = _lib.load()
    d, h, w, fz, fy, fx, kz, ky, kx, ic, oc, items = GEOMS[name]
    geo = UpconvGeom(d, h, w, fz, fy, fx, kz, ky, kx, ic, oc)
    gen = torch.Generator().manual_seed(len(name))
    wt = (torch.randn(oc, ic, kz, ky, kx, generator=gen) / np.sqrt(ic * kz * ky * kx)).cuda()
    bias = torch.randn(oc, generator=gen).cuda()
    src = torch.randn(items, d, h, w, ic, generator=gen).cuda()
    taps = (1 if kz == 1 else 2) * (1 if ky == 1 else 2) * (1 if kx == 1 else 2)
    n = fz * fy * fx * taps * ic * oc
    w_eff, w_eff_t, w_eff_k = (torch.empty(n, device="cuda") for _ in range(3))
    check(lib.rcb_fold_poly(ptr(wt), C.byref(geo), ptr(w_eff), ptr(w_eff_t), stream()))
    check(lib.rcb_fold_poly_k(ptr(wt), C.byref(geo), ptr(w_eff_k), stream()))
    out_shape = (items, d * fz, h * fy, w * fx, oc)
    ref = torch.zeros(out_shape, device="cuda")
    got = torch.full(out_shape, 3.0, device="cuda")
    check(lib.rcb_upconv_fwd(ptr(src), ptr(w_eff), ptr(bias), ptr(ref), C.byref(geo), items, 1, stream()))
    check(lib.rcb_upconv_fwd_tc(ptr(src), ptr(w_eff_k), ptr(bias), ptr(got), C.byref(geo), items, 1, stream()))
    torch.cuda.synchronize()
    scale = float(src.norm(dim=-1).max()) * float(wt.flatten(1).norm(dim=1).max())
    err = float((got - ref).abs().max())
    assert err < 2e-3 * scale, (err, scale)
    # data gradient (with the LeakyReLU mask of the producing stage)
    d_out = torch.randn(out_shape, generator=gen).cuda()
    act = torch.randn(items, d, h, w, ic, generator=gen).cuda()
    ref_b = torch.zeros_like(src)
    got_b = torch.full_like(src, 3.0)
    check(lib.rcb_upconv_bwd(ptr(d_out), ptr(w_eff_t), ptr(act), ptr(ref_b), C.byref(geo), items, stream()))
    check(lib.rcb_upconv_bwd_tc(ptr(d_out), ptr(w_eff), ptr(act), ptr(got_b), C.byref(geo), items, stream()))
    torch.cuda.synchronize()
    scale_b = float(d_out.norm(dim=-1).max()) * float(wt.norm()) / np.sqrt(ic) * 4
    err_b = float((got_b - ref_b).abs().max())
    assert err_b < 2e-3 * scale_b, (err_b, scale_b)
    print(f"[conv_tc {name}] fwd err {err:.2e} (scale {scale:.2f}); bwd err {err_b:.2e} (scale {scale_b:.2f})")


@pytest.mark.parametrize("name", ["cifar_conv2", "cifar_conv3", "conv1_poly_2d_wg"])
def test_upconv_wgrad_tc_matches_simt(name):
    """Weight gradient of the polyphase conv on tcgen05 (channel-major copies through rcb_transpose) against the
    SIMT split-K kernel: TF32 operands, stated tolerance 2e-3 of |src|.|d_out| summed over the K extent."""
    from recombiner_b200 import _lib
    from recombiner_b200._lib import UpconvGeom, check, ptr, stream
    lib = _lib.load()
    geoms = dict(GEOMS, conv1_poly_2d_wg=(1, 8, 16, 1, 4, 4, 1, 5, 5, 128, 64, 3))
    d, h, w, fz, fy, fx, kz, ky, kx, ic, oc, items = geoms[name]
    geo = UpconvGeom(d, h, w, fz, fy, fx, kz, ky, kx, ic, oc)
    gen = torch.Generator().manual_seed(7 + len(name))
    src = torch.randn(items, d, h, w, ic, generator=gen).cuda()
    d_out = torch.randn(items, d * fz, h * fy, w * fx, oc, generator=gen).cuda()
    n = fy * fx * 4 * ic * oc
    ref, got = torch.zeros(n, device="cuda"), torch.full((n,), 3.0, device="cuda")
    check(lib.rcb_upconv_wgrad(ptr(src), ptr(d_out), ptr(ref), C.byref(geo), items, stream()))
    rows_in, rows_out = items * h * w, items * h * fy * w * fx
    srcT, doutT = torch.empty(3, ic, rows_in, device="cuda"), torch.empty(oc, rows_out, device="cuda")
    check(lib.rcb_transpose_xshift(ptr(src), ptr(srcT), rows_in, ic, w, stream()))
    check(lib.rcb_transpose_phases(ptr(d_out), ptr(doutT), rows_out, oc, h, w, fy, fx, stream()))
    padded = torch.nn.functional.pad(src, (0, 0, 1, 1))                      # zeros left and right of every line
    for k in range(3):
        assert torch.equal(srcT[k], padded[:, :, :, k:k + w].reshape(rows_in, ic).t())
    planes = d_out.reshape(items, h, fy, w, fx, oc).permute(5, 0, 2, 4, 1, 3).reshape(oc, rows_out)
    assert torch.equal(doutT, planes)
    check(lib.rcb_upconv_wgrad_tc(ptr(srcT), ptr(doutT), ptr(got), C.byref(geo), items, stream()))
    torch.cuda.synchronize()
    scale = float(np.sqrt(rows_in)) * 3.0
    err = float((got - ref).abs().max())
    print(f"[wgrad_tc {name}] err {err:.2e} (max |ref| {float(ref.abs().max()):.2f}, scale {scale:.1f})")
    assert err < 2e-3 * scale * 4, (err, scale)


@pytest.mark.parametrize("h,w,items", [(16, 16, 5), (40, 44, 20)])
def test_upconv_fwd_fp16_operands_match_simt_on_rounded_inputs(h, w, items):
    """rcb_upconv_fwd_tc_h (fp16 activations and weights, fp32 accumulation) against the fp32 SIMT engine run on
    the same fp16-rounded activations: what is left is the fp16 rounding of the weights and the summation order."""
    from recombiner_b200 import _lib
    from recombiner_b200._lib import UpconvGeom, check, ptr, stream
    lib = _lib.load()
    ic, oc = 64, 16
    geo = UpconvGeom(1, h, w, 1, 2, 2, 1, 3, 3, ic, oc)
    gen = torch.Generator().manual_seed(h * w)
    wt = (torch.randn(oc, ic, 1, 3, 3, generator=gen) / np.sqrt(ic * 9)).cuda()
    bias = torch.randn(oc, generator=gen).cuda()
    src = torch.randn(items, 1, h, w, ic, generator=gen).cuda()
    n = 4 * 4 * ic * oc
    w_eff, w_eff_t, w_eff_k = (torch.empty(n, device="cuda") for _ in range(3))
    check(lib.rcb_fold_poly(ptr(wt), C.byref(geo), ptr(w_eff), ptr(w_eff_t), stream()))
    check(lib.rcb_fold_poly_k(ptr(wt), C.byref(geo), ptr(w_eff_k), stream()))
    src_h = torch.empty(src.shape, dtype=torch.float16, device="cuda")
    w_h = torch.empty(n, dtype=torch.float16, device="cuda")
    check(lib.rcb_to_half(ptr(src), ptr(src_h), src.numel(), stream()))
    check(lib.rcb_to_half(ptr(w_eff_k), ptr(w_h), n, stream()))
    assert torch.equal(src_h, src.half()) and torch.equal(w_h, w_eff_k.half())
    out_shape = (items, 1, 2 * h, 2 * w, oc)
    ref = torch.zeros(out_shape, device="cuda")
    got = torch.full(out_shape, 3.0, device="cuda")
    src_r = src_h.float().contiguous()
    check(lib.rcb_upconv_fwd(ptr(src_r), ptr(w_eff), ptr(bias), ptr(ref), C.byref(geo), items, 1, stream()))
    check(lib.rcb_upconv_fwd_tc_h(ptr(src_h), ptr(w_h), ptr(bias), ptr(got), C.byref(geo), items, 1, stream()))
    torch.cuda.synchronize()
    scale = float(src.norm(dim=-1).max()) * float(wt.flatten(1).norm(dim=1).max())
    err = float((got - ref).abs().max())
    assert err < 1e-3 * scale, (err, scale)


@pytest.mark.parametrize("h,w,items", [(16, 16, 5), (40, 44, 20)])
@pytest.mark.parametrize("act_kind", [0, 1, 2])
def test_upconv_bwd_f2_matches_simt(h, w, items, act_kind):
    """Resident-weight x2 data-gradient kernel against the fp32 SIMT engine; LeakyReLU mask absent, from fp32 and
    from fp16 activations (the mask only reads signs, so both must agree with the SIMT kernel on the rounded ones)."""
    from recombiner_b200 import _lib
    from recombiner_b200._lib import UpconvGeom, check, ptr, stream
    lib = _lib.load()
    ic, oc = 64, 16
    geo = UpconvGeom(1, h, w, 1, 2, 2, 1, 3, 3, ic, oc)
    gen = torch.Generator().manual_seed(h + w + act_kind)
    wt = (torch.randn(oc, ic, 1, 3, 3, generator=gen) / np.sqrt(ic * 9)).cuda()
    n = 4 * 4 * ic * oc
    w_eff, w_eff_t, w_bk = (torch.empty(n, device="cuda") for _ in range(3))
    check(lib.rcb_fold_poly(ptr(wt), C.byref(geo), ptr(w_eff), ptr(w_eff_t), stream()))
    check(lib.rcb_fold_poly_bwd_f2(ptr(w_eff), C.byref(geo), ptr(w_bk), stream()))
    d_out = torch.randn(items, 1, 2 * h, 2 * w, oc, generator=gen).cuda()
    act = torch.randn(items, 1, h, w, ic, generator=gen).cuda().half().float()
    act_arg = {0: None, 1: act, 2: act.half()}[act_kind]
    ref = torch.zeros_like(act)
    got = torch.full_like(act, 3.0)
    check(lib.rcb_upconv_bwd(ptr(d_out), ptr(w_eff_t), ptr(act if act_kind else None), ptr(ref), C.byref(geo), items, stream()))
    check(lib.rcb_upconv_bwd_f2(ptr(d_out), ptr(w_bk), ptr(act_arg), act_kind, ptr(got), C.byref(geo), items, stream()))
    torch.cuda.synchronize()
    scale = float(d_out.norm(dim=-1).max()) * float(wt.norm()) / np.sqrt(ic) * 4
    err = float((got - ref).abs().max())
    assert err < 2e-3 * scale, (err, scale)


@pytest.mark.parametrize("h,w,items", [(16, 16, 5), (40, 44, 20)])
def test_upconv_bwd_f2_fp16_output_is_the_scaled_rounded_fp32_output(h, w, items):
    """rcb_upconv_bwd_f2_oh = fp16(out_scale * rcb_upconv_bwd_f2), same accumulators: equal up to the last fp16 bit
    (the scale is a power of two), and clamped instead of overflowing."""
    from recombiner_b200 import _lib
    from recombiner_b200._lib import UpconvGeom, check, ptr, stream
    lib = _lib.load()
    ic, oc = 64, 16
    geo = UpconvGeom(1, h, w, 1, 2, 2, 1, 3, 3, ic, oc)
    gen = torch.Generator().manual_seed(h + w)
    wt = (torch.randn(oc, ic, 1, 3, 3, generator=gen) / np.sqrt(ic * 9)).cuda()
    n = 4 * 4 * ic * oc
    w_eff, w_eff_t, w_bk = (torch.empty(n, device="cuda") for _ in range(3))
    check(lib.rcb_fold_poly(ptr(wt), C.byref(geo), ptr(w_eff), ptr(w_eff_t), stream()))
    check(lib.rcb_fold_poly_bwd_f2(ptr(w_eff), C.byref(geo), ptr(w_bk), stream()))
    d_out = (torch.randn(items, 1, 2 * h, 2 * w, oc, generator=gen) * 1e-4).cuda()
    act = torch.randn(items, 1, h, w, ic, generator=gen).cuda().half()
    ref = torch.zeros(act.shape, device="cuda")
    got = torch.full(act.shape, 3.0, dtype=torch.float16, device="cuda")
    check(lib.rcb_upconv_bwd_f2(ptr(d_out), ptr(w_bk), ptr(act), 2, ptr(ref), C.byref(geo), items, stream()))
    check(lib.rcb_upconv_bwd_f2_oh(ptr(d_out), ptr(w_bk), ptr(act), 2, ptr(got), 8192.0, C.byref(geo), items, stream()))
    torch.cuda.synchronize()
    assert torch.equal(got, (ref * 8192.0).half())
    check(lib.rcb_upconv_bwd_f2_oh(ptr(d_out), ptr(w_bk), ptr(act), 2, ptr(got), 2.0 ** 40, C.byref(geo), items, stream()))
    torch.cuda.synchronize()
    assert torch.isfinite(got.float()).all() and float(got.float().abs().max()) == 65504.0


@pytest.mark.parametrize("h,w,items", [(16, 16, 5), (40, 44, 20)])
def test_upconv_bwd_f2_fp16_in_fp16_out_matches_simt_on_rounded_inputs(h, w, items):
    """rcb_upconv_bwd_f2_hh (fp16 gradient in, fp16 weights, one kind::f16 MMA per product) against the fp32 SIMT
    engine on the same fp16-rounded gradient: weight rounding, summation order and the fp16 rounding of the output."""
    from recombiner_b200 import _lib
    from recombiner_b200._lib import UpconvGeom, check, ptr, stream
    lib = _lib.load()
    ic, oc = 64, 16
    geo = UpconvGeom(1, h, w, 1, 2, 2, 1, 3, 3, ic, oc)
    gen = torch.Generator().manual_seed(h * w + 1)
    wt = (torch.randn(oc, ic, 1, 3, 3, generator=gen) / np.sqrt(ic * 9)).cuda()
    n = 4 * 4 * ic * oc
    w_eff, w_eff_t, w_bk = (torch.empty(n, device="cuda") for _ in range(3))
    check(lib.rcb_fold_poly(ptr(wt), C.byref(geo), ptr(w_eff), ptr(w_eff_t), stream()))
    check(lib.rcb_fold_poly_bwd_f2(ptr(w_eff), C.byref(geo), ptr(w_bk), stream()))
    w_bk_h = torch.empty(n, dtype=torch.float16, device="cuda")
    check(lib.rcb_to_half(ptr(w_bk), ptr(w_bk_h), n, stream()))
    d_h = torch.randn(items, 1, 2 * h, 2 * w, oc, generator=gen).cuda().half()
    d_r = d_h.float().contiguous()
    act = torch.randn(items, 1, h, w, ic, generator=gen).cuda().half()
    act_f = act.float().contiguous()
    ref = torch.zeros(act.shape, device="cuda")
    got = torch.full(act.shape, 3.0, dtype=torch.float16, device="cuda")
    check(lib.rcb_upconv_bwd(ptr(d_r), ptr(w_eff_t), ptr(act_f), ptr(ref), C.byref(geo), items, stream()))
    check(lib.rcb_upconv_bwd_f2_hh(ptr(d_h), ptr(w_bk_h), ptr(act), 2, ptr(got), 4.0, C.byref(geo), items, stream()))
    torch.cuda.synchronize()
    bound = float(d_r.norm(dim=-1).max()) * float(wt.norm()) / np.sqrt(ic) * 4
    err = float((got.float() / 4.0 - ref).abs().max())
    print(f"[bwd_f2_hh {h}x{w} x{items}] err {err:.2e} bound {bound:.2e} max|ref| {float(ref.abs().max()):.2e}")
    assert err < 1e-3 * bound, (err, bound)


@pytest.mark.parametrize("h,w,items", [(8, 8, 7), (8, 8, 701), (16, 16, 3), (40, 44, 5)])
@pytest.mark.parametrize("masked", [False, True])
def test_upconv_bwd_f2w_matches_simt_on_rounded_inputs(h, w, items, masked):
    """Resident-weight fp16 data gradient of the 64 -> 64 x2 stage against the fp32 SIMT engine run on the same
    fp16-rounded gradient: what is left is the fp16 rounding of the weights and the summation order.  Two 8 x 8
    items per tile (odd count; more than two tiles per CTA) and one item per tile (ragged)."""
    from recombiner_b200 import _lib
    from recombiner_b200._lib import UpconvGeom, check, ptr, stream
    lib = _lib.load()
    ic = oc = 64
    geo = UpconvGeom(1, h, w, 1, 2, 2, 1, 3, 3, ic, oc)
    assert lib.rcb_upconv_bwd_f2w_eligible(C.byref(geo)) == 1
    gen = torch.Generator().manual_seed(h + w + items)
    wt = (torch.randn(oc, ic, 1, 3, 3, generator=gen) / np.sqrt(ic * 9)).cuda()
    n = 4 * 4 * ic * oc
    w_eff, w_eff_t, w_bk = (torch.empty(n, device="cuda") for _ in range(3))
    check(lib.rcb_fold_poly(ptr(wt), C.byref(geo), ptr(w_eff), ptr(w_eff_t), stream()))
    check(lib.rcb_fold_poly_bwd_f2w(ptr(w_eff), C.byref(geo), ptr(w_bk), stream()))
    w_bk_h = torch.empty(n, dtype=torch.float16, device="cuda")
    check(lib.rcb_to_half(ptr(w_bk), ptr(w_bk_h), n, stream()))
    scale = 256.0
    d_true = torch.randn(items, 1, 2 * h, 2 * w, oc, generator=gen).cuda() / scale
    d_h = (d_true * scale).half()
    d_r = (d_h.float() / scale).contiguous()
    act = torch.randn(items, 1, h, w, ic, generator=gen).cuda().half()
    act_f = act.float().contiguous()
    ref = torch.zeros(act.shape, device="cuda")
    got = torch.full(act.shape, 3.0, device="cuda")
    check(lib.rcb_upconv_bwd(ptr(d_r), ptr(w_eff_t), ptr(act_f if masked else None), ptr(ref), C.byref(geo), items, stream()))
    check(lib.rcb_upconv_bwd_f2w(ptr(d_h), ptr(w_bk_h), ptr(act if masked else None), ptr(got), 1.0 / scale, C.byref(geo),
                                 items, stream()))
    torch.cuda.synchronize()
    bound = float(d_r.norm(dim=-1).max()) * float(wt.norm()) / np.sqrt(ic) * 4
    err = float((got - ref).abs().max())
    print(f"[bwd_f2w {h}x{w} x{items}] err {err:.2e} bound {bound:.2e} max|ref| {float(ref.abs().max()):.2e}")
    assert err < 2e-4 * bound, (err, bound)
    # fp16 output = the fp32 output (same accumulators, same power-of-two factor) rounded once
    got_h = torch.full(act.shape, 3.0, dtype=torch.float16, device="cuda")
    check(lib.rcb_upconv_bwd_f2w_oh(ptr(d_h), ptr(w_bk_h), ptr(act if masked else None), ptr(got_h), 1.0 / scale, C.byref(geo),
                                    items, stream()))
    torch.cuda.synchronize()
    assert torch.equal(got_h, got.half())


@pytest.mark.parametrize("name", ["cifar_conv2", "wide_2d_ragged", "video_conv2_3d", "protein_conv2", "f2w_ragged", "f2w_8x8_many"])
def test_upconv_fwd_fp16_in_fp16_out_general_kernel(name):
    """rcb_upconv_fwd_tc_hh (general kernel, kind::f16 MMAs, fp16 result) against the SIMT engine on the rounded inputs."""
    from recombiner_b200 import _lib
    from recombiner_b200._lib import UpconvGeom, check, ptr, stream
    lib = _lib.load()
    d, h, w, fz, fy, fx, kz, ky, kx, ic, oc, items = GEOMS[name]
    geo = UpconvGeom(d, h, w, fz, fy, fx, kz, ky, kx, ic, oc)
    gen = torch.Generator().manual_seed(len(name) + 7)
    wt = (torch.randn(oc, ic, kz, ky, kx, generator=gen) / np.sqrt(ic * kz * ky * kx)).cuda()
    bias = torch.randn(oc, generator=gen).cuda()
    src_h = torch.randn(items, d, h, w, ic, generator=gen).cuda().half()
    taps = (1 if kz == 1 else 2) * (1 if ky == 1 else 2) * (1 if kx == 1 else 2)
    n = fz * fy * fx * taps * ic * oc
    w_eff, w_eff_t, w_eff_k = (torch.empty(n, device="cuda") for _ in range(3))
    check(lib.rcb_fold_poly(ptr(wt), C.byref(geo), ptr(w_eff), ptr(w_eff_t), stream()))
    check(lib.rcb_fold_poly_k(ptr(wt), C.byref(geo), ptr(w_eff_k), stream()))
    w_h = w_eff_k.half()
    out_shape = (items, d * fz, h * fy, w * fx, oc)
    ref = torch.zeros(out_shape, device="cuda")
    got = torch.full(out_shape, 3.0, device="cuda", dtype=torch.float16)
    src_r = src_h.float().contiguous()
    check(lib.rcb_upconv_fwd(ptr(src_r), ptr(w_eff), ptr(bias), ptr(ref), C.byref(geo), items, 1, stream()))
    check(lib.rcb_upconv_fwd_tc_hh(ptr(src_h), ptr(w_h), ptr(bias), ptr(got), C.byref(geo), items, 1, stream()))
    torch.cuda.synchronize()
    scale = float(src_r.norm(dim=-1).max()) * float(wt.flatten(1).norm(dim=1).max())
    err = float((got.float() - ref).abs().max())
    assert err < 1e-3 * scale + 1e-3 * float(ref.abs().max()), (err, scale)      # + fp16 rounding of the result


def test_gemm_tc_fp16_output_is_the_rounded_fp32_output():
    from recombiner_b200 import _lib
    from recombiner_b200._lib import check, ptr, stream
    lib = _lib.load()
    M, N, K = 300, 4096, 2048
    gen = torch.Generator().manual_seed(3)
    A = torch.randn(M, K, generator=gen).cuda()
    Bt = (torch.randn(N, K, generator=gen) / np.sqrt(K)).cuda()
    bias = torch.randn(64, generator=gen).cuda()
    c32 = torch.zeros(M, N, device="cuda")
    c16 = torch.zeros(M, N, device="cuda", dtype=torch.float16)
    check(lib.rcb_gemm_tc(ptr(A), K, ptr(Bt), K, ptr(c32), N, M, N, K, ptr(bias), 64, 1, 0, stream()))
    check(lib.rcb_gemm_tc_oh(ptr(A), K, ptr(Bt), K, ptr(c16), N, M, N, K, ptr(bias), 64, 1, stream()))
    torch.cuda.synchronize()
    assert torch.equal(c16, c32.half())


def test_gemm_tc_fp16_operands():
    """rcb_gemm_tc_hh (kind::f16) against an fp32 matmul of the same fp16-rounded operands."""
    from recombiner_b200 import _lib
    from recombiner_b200._lib import check, ptr, stream
    lib = _lib.load()
    M, N, K = 300, 4096, 2048
    gen = torch.Generator().manual_seed(5)
    A = torch.randn(M, K, generator=gen).cuda().half()
    Bt = (torch.randn(N, K, generator=gen) / np.sqrt(K)).cuda().half()
    bias = torch.randn(64, generator=gen).cuda()
    c16 = torch.zeros(M, N, device="cuda", dtype=torch.float16)
    check(lib.rcb_gemm_tc_hh(ptr(A), K, ptr(Bt), K, ptr(c16), N, M, N, K, ptr(bias), 64, 1, stream()))
    torch.cuda.synchronize()
    ref = torch.nn.functional.leaky_relu(A.float() @ Bt.float().t() + bias.repeat(N // 64), 0.01)
    err = float((c16.float() - ref).abs().max())
    assert err < 2e-3 * float(ref.abs().max()), err
