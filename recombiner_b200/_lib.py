"""ctypes binding of librecombiner_b200.so (the C ABI in include/recombiner_b200.h).

There is no CPU fallback: importing the kernels without the built library, or
calling them without a CUDA device, raises.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "librecombiner_b200.so")

P = C.c_void_p
I32 = C.c_int
I64 = C.c_int64
F32 = C.c_float
F64 = C.c_double


class SampleArgs(C.Structure):
    _fields_ = [(n, P) for n in ("loc", "log_scale", "mask", "sample", "g2p", "perm", "row_map",
                                 "eps_w", "eps_l", "hw", "lpe", "lpe_slot", "eps_w_store", "eps_l_store")] + \
               [("seed", I64), ("row_offset", I64)] + \
               [(n, I32) for n in ("rows", "S", "P", "n_w", "n_l", "ld_hw", "step", "tensor_id", "accumulate",
                                   "rows_per_datum", "sp_total", "lpe_c")] + [("dyn", P), ("lpe_h", P), ("hw_h", P), ("p2g", P), ("fast_math", I32)]


class UpconvGeom(C.Structure):
    _fields_ = [(n, I32) for n in ("d", "h", "w", "fz", "fy", "fx", "kz", "ky", "kx", "ic", "oc")]


class MlpArgs(C.Structure):
    _fields_ = [(n, P) for n in ("wt", "xt", "pe", "y", "dy", "y_pred", "d_pe", "d_wt", "sqerr", "pe_base")] + \
               [("x_row_stride", I64), ("pitch_z", I64), ("pitch_y", I64)] + \
               [(n, I32) for n in ("items", "S", "pix", "n_f", "out", "ld_w", "mode", "ph", "pw")] + \
               [("coef", F32), ("w0", F32), ("d_wt_h_scale", F32), ("ld_wh", I32), ("d_wt_h", P), ("pe_half", I32),
                ("x_tab", P), ("x_axes", I32), ("x_nfreq", I32), ("x_size", I32 * 3), ("x_off", I32 * 3), ("d_pe_h", P)]


class UpdateArgs(C.Structure):
    _fields_ = [(n, P) for n in ("loc", "log_scale", "mask", "p_loc", "p_log_scale", "beta", "group_idx", "p2g",
                                 "perm_inv", "row_children", "d_hw", "d_lpe", "eps_w", "eps_l", "lpe_slot", "g_loc",
                                 "g_log_scale", "m1_loc", "v_loc", "m1_ls", "v_ls", "kl_out")] + \
               [("seed", I64), ("row_offset", I64)] + \
               [(n, I32) for n in ("src_rows", "rows", "n_children", "S", "P", "n_w", "n_l", "ld_hw", "G",
                                   "step", "tensor_id", "adam", "p_scale_direct", "rows_per_datum", "sp_total", "lpe_c")] + \
               [(n, F32) for n in ("adam_step_size", "adam_bc2_sqrt", "b1", "b2", "adam_eps", "beta_scalar",
                                   "grad_scale")] + [("dyn", P)] + \
               [(n, P) for n in ("red_mu", "red_sig", "red_mu_l", "red_sig_l")] + [("fast_math", I32)]


class ReduceArgs(C.Structure):
    _fields_ = [("d_hw", P), ("d_lpe", P), ("eps_w", P * 3), ("eps_l", P), ("lpe_slot", P), ("red_mu", P),
                ("red_sig", P * 3), ("red_mu_l", P), ("red_sig_l", P)] + \
               [(n, I32) for n in ("rows", "S", "n_w", "n_l", "ld_hw", "n_levels", "rows_per_datum", "sp_total", "lpe_c")]


class RecArgs(C.Structure):
    _fields_ = [(n, P) for n in ("pair_row", "pair_block", "q_loc", "q_scale", "p_loc", "p_scale", "group_start",
                                 "group_end", "tables", "gumbel", "idx_out", "z_out", "logw_out", "sample", "mask",
                                 "beta", "coded")] + \
               [(n, I32) for n in ("n_pairs", "P", "G", "n_cand", "max_D", "apply")] + \
               [("workspace", P), ("workspace_bytes", I64)]


class StepState(C.Structure):
    _fields_ = [("seed", I64), ("step", I32), ("adam_step_size", F32), ("adam_bc2_sqrt", F32), ("beta_scalar", F32)]


STRUCTS = {"rcb_step_state": StepState, "rcb_sample_args": SampleArgs, "rcb_upconv_geom": UpconvGeom, "rcb_mlp_args": MlpArgs,
           "rcb_update_args": UpdateArgs, "rcb_rec_args": RecArgs, "rcb_reduce_args": ReduceArgs}

# name -> argtypes (restype is int unless listed in _RESTYPES)
SIGNATURES = {
    "rcb_version": [],
    "rcb_last_error": [],
    "rcb_set_step_state": [P, I64, I32, F32, F32, F32, P],
    "rcb_adam_flat": [P, P, P, P, I64, F32, F32, F32, F32, F32, P, P],
    "rcb_step_stats": [P, I32, F64, P, P, P],
    "rcb_fit_sample": [C.POINTER(SampleArgs), P],
    "rcb_gemm": [P, I32, P, I32, P, I32, I32, I32, I32, P, I32, I32, I32, I32, P],
    "rcb_gemm_tc": [P, I32, P, I32, P, I32, I32, I32, I32, P, I32, I32, I32, P],
    "rcb_fold_poly": [P, C.POINTER(UpconvGeom), P, P, P],
    "rcb_fold_dense": [P, C.POINTER(UpconvGeom), P, P, P],
    "rcb_fold_poly_k": [P, C.POINTER(UpconvGeom), P, P],
    "rcb_upconv_fwd_tc": [P, P, P, P, C.POINTER(UpconvGeom), I32, I32, P],
    "rcb_upconv_fwd_tc_h": [P, P, P, P, C.POINTER(UpconvGeom), I32, I32, P],
    "rcb_upconv_fwd_tc_hh": [P, P, P, P, C.POINTER(UpconvGeom), I32, I32, P],
    "rcb_gemm_tc_oh": [P, I32, P, I32, P, I32, I32, I32, I32, P, I32, I32, P],
    "rcb_gemm_tc_hh": [P, I32, P, I32, P, I32, I32, I32, I32, P, I32, I32, P],
    "rcb_gemm_tc_batch": [I32, P, I32, P, P, P, I32, I32, P, P, I32, F32, P],
    "rcb_gemm_tc_h": [P, I32, P, I32, P, I32, I32, I32, I32, P, I32, I32, I32, P],
    "rcb_upconv_fwd_tc_oh": [P, P, P, P, C.POINTER(UpconvGeom), I32, I32, P],
    "rcb_upconv_bwd_tc_ah": [P, P, P, P, C.POINTER(UpconvGeom), I32, P],
    "rcb_fold_poly_bwd_f2": [P, C.POINTER(UpconvGeom), P, P],
    "rcb_upconv_bwd_f2": [P, P, P, I32, P, C.POINTER(UpconvGeom), I32, P],
    "rcb_upconv_bwd_f2_oh": [P, P, P, I32, P, F32, C.POINTER(UpconvGeom), I32, P],
    "rcb_upconv_bwd_f2_hh": [P, P, P, I32, P, F32, C.POINTER(UpconvGeom), I32, P],
    "rcb_upconv_bwd_f2w_oh": [P, P, P, P, F32, C.POINTER(UpconvGeom), I32, P],
    "rcb_upconv_bwd_f2w_eligible": [C.POINTER(UpconvGeom)],
    "rcb_fold_poly_bwd_f2w": [P, C.POINTER(UpconvGeom), P, P],
    "rcb_upconv_bwd_f2w": [P, P, P, P, F32, C.POINTER(UpconvGeom), I32, P],
    "rcb_to_half": [P, P, I64, P],
    "rcb_upconv_bwd_tc": [P, P, P, P, C.POINTER(UpconvGeom), I32, P],
    "rcb_upconv_fwd": [P, P, P, P, C.POINTER(UpconvGeom), I32, I32, P],
    "rcb_upconv_bwd": [P, P, P, P, C.POINTER(UpconvGeom), I32, P],
    "rcb_upconv_wgrad": [P, P, P, C.POINTER(UpconvGeom), I32, P],
    "rcb_upconv_wgrad_tc": [P, P, P, C.POINTER(UpconvGeom), I32, P],
    "rcb_unfold_poly": [P, C.POINTER(UpconvGeom), P, P],
    "rcb_unfold_dense": [P, C.POINTER(UpconvGeom), P, P],
    "rcb_colsum": [P, I64, I32, I32, P, P],
    "rcb_mlp": [C.POINTER(MlpArgs), P],
    "rcb_mlp_tc": [C.POINTER(MlpArgs), P],
    "rcb_fourier_table": [P, I32, C.POINTER(F32), I32, P],
    "rcb_transpose": [P, I64, P, I64, I32, I32, P],
    "rcb_transpose_phases": [P, P, I64, I32, I32, I32, I32, I32, P],
    "rcb_transpose_xshift": [P, P, I64, I32, I32, P],
    "rcb_fit_update": [C.POINTER(UpdateArgs), P],
    "rcb_fit_reduce": [C.POINTER(ReduceArgs), P],
    "rcb_std_transform": [P, P, I64, P],
    "rcb_group_kl": [P, P, P, P, P, P, P, I32, I32, I32, P],
    "rcb_anneal_beta": [P, P, P, I32, I32, F64, F64, F64, F64, P],
    "rcb_pick_block": [P, P, P, I32, I32, P],
    "rcb_rec_table": [P, P, P, I32, I32, P],
    "rcb_rec_encode": [C.POINTER(RecArgs), P],
    "rcb_rec_order": [P, P, P, I32, I32, P],
    "rcb_ubench_dfma": [P, I32, I32, C.POINTER(F64), P],
    "rcb_rec_decode": [P, P, P, P, P, P, P, P, P, P, I32, I32, I32, P],
    "rcb_prior_suffstats": [P, P, P, I32, I32, P],
    "rcb_prior_from_stats": [P, P, P, I64, I32, P],
}
_RESTYPES = {"rcb_last_error": C.c_char_p}

_lib = None


class KernelError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library (building nothing: see recombiner_b200._build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KernelError(
            f"{LIB_PATH} is missing: build it with `python -m recombiner_b200._build` "
            "(the sm_100a CUDA extension is the only implementation; there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


COUNTERS = {"launches": 0}


def check(rc: int, what: str = "") -> None:
    """Every C-ABI call launches exactly one kernel of ours and goes through here."""
    COUNTERS["launches"] += 1
    if rc != 0:
        msg = load().rcb_last_error().decode("utf-8", "replace")
        raise KernelError(f"{what or 'librecombiner_b200'} failed (rc={rc}): {msg}")


def ptr(t) -> int:
    """Device pointer of a torch tensor (None -> NULL).  The kernels launch on the CURRENT device and stream, so a
    tensor that lives on another GPU is refused instead of being dereferenced there."""
    if t is None:
        return None
    if not t.is_cuda:
        raise KernelError("librecombiner_b200 kernels take CUDA tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise KernelError("non-contiguous tensor passed to a kernel")
    import torch
    if t.device.index != torch.cuda.current_device():
        raise KernelError(f"tensor on {t.device} passed to a kernel launching on cuda:{torch.cuda.current_device()} "
                          "(models select their device with torch.cuda.set_device at construction)")
    return t.data_ptr()


def use_device(device) -> None:
    """Make `device` the current CUDA device: every C-ABI call launches on the current device and stream."""
    import torch
    dev = torch.device(device)
    if dev.type == "cuda" and dev.index is not None and dev.index != torch.cuda.current_device():
        torch.cuda.set_device(dev)


def stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
