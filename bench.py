#!/usr/bin/env python
"""Benchmark of the RECOMBINER hot path (BASELINE.json metric: datapoints compressed/s,
CIFAR-10-shape 32x32; secondary: REC candidates/s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the UNMODIFIED reference (oracle/_ref) on the host cores
    python bench.py --workload kodak|audio|video|protein      # one fit step of another modality shape (step ms + roofline)

Workload (configs[1] of BASELINE.json): 1024 synthetic CIFAR-shape images per GPU,
S=5 MC samples, random-init prior (mu_p=0, sigma_p=softplus(-2)/6), synthetic
grouping at 0.52 bpp (G blocks of 16 bits).  One *step* = one pass of the fit hot
path over the batch (sample -> reparam GEMMs -> folded upsampler -> fused SIREN MLP
fwd+loss+bwd -> conv/reparam data-gradients -> KL-gradient + Adam, beta annealing
every 10th step), i.e. one iteration of test_model.py:622-635.  One *REC round* codes
one block of every row (test_model.py:806-818).

A full compression under the reference schedule (main_compression.py:148-162) is
`steps_full = 30000 + G*max(30000//G, 50)` steps and G rounds, so

    value = datapoints / (steps_full * t_step + G * t_round)      [datapoints compressed/s]

with t_step and t_round both measured live (CUDA events, K timed iterations each,
after W warm-ups).  The working set of a step (~2 GB) exceeds L2 (126 MB).

Beside that extrapolation both arms run `e2e_short`: ONE short schedule to completion through the public calls
(`optimize_posteriors` + `compress_posteriors`, main_compression.py:148-162) on the same 8 rows, timed by wall clock
with host inputs, reporting PSNR and bpp -- a directly measured compression.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS_PER_GPU = 1024
S = 5
TOTAL_BITS = 532.0          # -> G = 33..34 blocks, 0.52 bpp on 1024 pixels
N_CAND = 65536


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full
# captures (profiles/README.md); None where no capture exists yet
TRAFFIC = {"mlp_fwd_bwd": 406.8e6, "update": 313.2e6, "sample": 116.4e6,     # profiles/r2_launches_final.csv (ncu, dram bytes per launch)
           "conv3_fwd": 288.9e6, "conv2_fwd": 152.4e6, "conv3_bwd": 471.7e6, "conv2_bwd": 256.4e6}
REC_TRAFFIC = {"picked": 36.6e6, "spread": 1672.9e6}                       # profiles/r2_ncu_rec_encode_staged.csv


def conv_bytes(eng, items: int):
    """Algorithmic HBM bytes per launch of the upsampler kernels (cifar shapes): every tensor read or written once."""
    g1, g2, g3 = eng.geoms
    n0 = g1.h * g1.w * g1.ic
    n1 = g2.h * g2.w * g2.ic
    n2 = g3.h * g3.w * g3.ic
    n3 = g3.h * g3.fy * g3.w * g3.fx * g3.oc
    h = 2 if (eng.half_acts and eng.f2_half) else 4          # bytes per stored activation element
    hp = 2 if (h == 2 and eng.half_pe and eng.tc_mlp and eng.n_f == 16) else 4      # positional encodings
    hg = 2 if (h == 2 and getattr(eng, "b2w", False) and eng.tc_mlp) else 4          # gradients between the MLP and conv2
    hd = 2 if (hg == 2 and eng.dense1 and getattr(eng, "M1_h", None) is not None) else 4     # gradient into the dense first stage
    per_item = {"conv2_fwd": n1 * h + n2 * h, "conv3_fwd": n2 * h + n3 * hp,
                # data gradients: with the fp16 exchange (engine.b2w) d_pe and d_a2 travel as fp16, d_a1 stays fp32
                "conv3_bwd": n3 * hg + n2 * h + n2 * hg, "conv2_bwd": n2 * hg + n1 * h + n1 * hd,
                "conv1_bwd": n1 * hd + n0 * 4}
    return {k: float(v) * items for k, v in per_item.items()}


def schedule(G: int):
    return 30000 + G * max(30000 // G, 50)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops_sustained"], tf_burst=d["bf16_tflops"], src="measured")
    return dict(hbm=6650.0, tf=1400.0, tf_burst=1590.0, src="fallback")


def make_workload(rows: int, seed: int, dataset: str = "cifar"):
    """Synthetic batch of a modality's shape with a random-init prior.  Plain tensors
    only (no model classes of either arm): targets, Fourier inputs, mapping weights drawn from one seeded generator,
    synthetic per-parameter KL bits for the grouping."""
    from recombiner_b200.config import configs          # the config dict equals the reference's (tests/golden/config_snapshot.json)
    cfg = configs[dataset]
    g = torch.Generator().manual_seed(seed)
    dd = cfg["data_dim"]
    axes = [-1 + 2 * ((0.5 + torch.arange(s)) / s) for s in cfg["pixel_sizes"]]
    coords = torch.stack(torch.meshgrid(*axes, indexing="ij"), -1).view(-1, dd)
    w = torch.exp(torch.linspace(0, np.log(1024), cfg["fourier_dim"] // (2 * dd)))
    arg = torch.matmul(coords.unsqueeze(-1), w.unsqueeze(0)).view(coords.shape[0], -1)
    x1 = torch.cat([torch.cos(np.pi * arg), torch.sin(np.pi * arg)], -1)                  # data/image.py:24-27
    y = torch.rand(rows, x1.shape[0], cfg["output_dim"], generator=g)
    dims = [cfg["input_dim"]] + cfg["hidden_dims"] + [cfg["output_dim"]]
    counts = [dims[i + 1] * (dims[i] + 1) for i in range(4)]
    W = sum(counts)
    L = int(np.prod([p // u for p, u in zip(cfg["pixel_sizes"], cfg["upsample_factors"])])) * cfg["latent_dim"]
    P = W + L
    bits = np.random.RandomState(0).gamma(2.0, 1.0, P)
    bits *= TOTAL_BITS / bits.sum()
    gm = torch.Generator().manual_seed(42)
    A = [(torch.rand(n, n, generator=gm) * 2 - 1) / n for n in counts]                      # prior_model.py:19-21
    up = {}
    for i, (ci, co, k) in enumerate([(128, 64, 5), (64, 64, 3), (64, 16, 3)], start=1):     # torch's default conv init
        bound = 1.0 / math.sqrt(ci * k ** dd)
        up[f"conv{i}.weight"] = (torch.rand(co, ci, *([k] * dd), generator=gm) * 2 - 1) * bound
        up[f"conv{i}.bias"] = (torch.rand(co, generator=gm) * 2 - 1) * bound
    return dict(cfg=cfg, dataset=dataset, dims=dims, rows=rows, x=x1[None].expand(rows, -1, -1), y=y, A=A, up=up, P=P, W=W,
                L=L, bits=bits, p_loc=torch.zeros(P), p_log_scale=torch.full((P,), -2.0))


def product_grouping(wl):
    from recombiner_b200.prior_model import get_grouping_by_kl
    if "grouping" not in wl:
        wl["grouping"] = get_grouping_by_kl(wl["bits"])
        wl["G"] = wl["grouping"][5]
    return wl["grouping"]


def build_model(wl, device, row_offset=0, rows=None, precision=None):
    """This repo's TestBNNmodel for the workload (single-level modalities)."""
    from recombiner_b200.prior_model import LinearTransform, Upsample
    from recombiner_b200.test_model import TestBNNmodel
    cfg = wl["cfg"]
    gi, gs, ge, g2p, p2g, G = product_grouping(wl)[:6]
    lt = LinearTransform(wl["dims"])
    with torch.no_grad():
        for p_, a_ in zip(lt.A, wl["A"]):
            p_.copy_(a_)
    up = Upsample(cfg["data_dim"], cfg["paddings"], cfg["layerwise_scale_factors"])
    up.load_state_dict({k: v.clone() for k, v in wl["up"].items()})
    kw = {}
    if cfg["patch"]:                                     # levels 2 / 3 of the weight hierarchy: same recipe, a quarter of the bits
        from recombiner_b200.prior_model import get_grouping_by_kl
        bw = np.random.RandomState(1).gamma(2.0, 1.0, wl["W"])
        g2 = get_grouping_by_kl(bw * (TOTAL_BITS / 4 / bw.sum()))
        for pre in ("h_", "hh_"):
            kw.update({pre + "p_loc": torch.zeros(wl["W"]), pre + "p_log_scale": torch.full((wl["W"],), -2.0),
                       pre + "init_log_scale": -4.0, pre + "param_to_group": g2[4], pre + "group_to_param": g2[3],
                       pre + "n_groups": g2[5], pre + "group_start_index": g2[1], pre + "group_end_index": g2[2],
                       pre + "group_idx": g2[0]})
    with contextlib.redirect_stdout(io.StringIO()):
        return TestBNNmodel(**kw, in_dim=cfg["input_dim"], hidden_dims=cfg["hidden_dims"], out_dim=cfg["output_dim"],
                            number_of_datapoints=rows or wl["rows"], upsample_factors=cfg["upsample_factors"],
                            latent_dim=cfg["latent_dim"], data_dim=cfg["data_dim"], pixel_sizes=cfg["pixel_sizes"],
                            patch=cfg["patch"], patch_nums=cfg["patch_nums"],
                            hierarchical_patch_nums=cfg["hierarchical_patch_nums"], dataset=wl["dataset"],
                            linear_transform=lt.to(device), upsample_net=up.to(device),
                            p_loc=wl["p_loc"], p_log_scale=wl["p_log_scale"], init_log_scale=-4.0,
                            param_to_group=p2g, group_to_param=g2p, n_groups=G, group_start_index=gs, group_end_index=ge,
                            group_idx=gi, device=device, random_seed=42, initial_beta=1e-8, row_offset=row_offset,
                            layer_scales=cfg["layerwise_scale_factors"], paddings=cfg["paddings"], precision=precision)


def oracle_case(wl):
    """The same workload in the oracle port's input format (fallback CPU arm when oracle/_ref did not travel)."""
    from oracle import cases
    from oracle.recombiner_oracle import grouping_by_kl
    rows, P = wl["rows"], wl["P"]
    gi, gs, ge, g2p, p2g, G = grouping_by_kl(wl["bits"])[:6]
    lvl = dict(loc=torch.zeros(rows, P), log_scale=torch.full((rows, P), -4.0), p_loc=wl["p_loc"],
               p_log_scale=wl["p_log_scale"], group_idx=gi, group_start=gs, group_end=ge, group_to_param=g2p,
               param_to_group=p2g, n_groups=G, coded=np.zeros((rows, G), dtype=bool), mask=torch.zeros(rows, P),
               sample=torch.zeros(rows, P), beta=torch.full((rows, G), 1e-8))
    return dict(shape=cases.shape_of("cifar"), rows=rows, A=wl["A"], w_up=wl["up"], x=wl["x"].contiguous(), y=wl["y"],
                lvl1=lvl)


def psnr8(y, y_hat):
    """Mean over rows of the 8-bit-rounded PSNR (utils.py:245-254), numpy only."""
    y, y_hat = np.asarray(y), np.asarray(y_hat)
    rec = np.round(np.clip(y_hat, 0, 1) * 255) / 255
    mse = np.mean((y.reshape(y.shape[0], -1) - rec.reshape(y.shape[0], -1)) ** 2, -1)
    return float(np.mean(20 * np.log10(1.0 / np.sqrt(mse))))


SHORT = dict(rows=8, n_fit=100, n_finetune=2)        # the e2e_short schedule, identical on both arms


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------- #
# algorithmic work per launch (DESIGN.md "Kernels"): MACs actually executed
# ----------------------------------------------------------------------------- #
def section_flops(items: int):
    pix, hid, fin, out = 1024, 32, 32, 3
    mlp_fwd = pix * (fin * hid + 2 * hid * hid + hid * out)
    mlp_dx = pix * (hid * out + 2 * hid * hid + 16 * hid)          # dx0 only for the 16 pe inputs
    conv2 = 256 * 64 * 256                                          # out px * oc * (4 taps * 64 ic)
    conv3 = 1024 * 16 * 256
    conv1 = 512 * 4096                                              # dense fold of up x4 + conv k5 on 2x2
    rep = 3 * 1056 * 1056 + 99 * 99
    macs = {"mlp_fwd_bwd": 2 * mlp_fwd + mlp_dx, "mlp_fwd": mlp_fwd, "conv1_fwd": conv1, "conv1_bwd": conv1,
            "conv2_fwd": conv2, "conv2_bwd": conv2, "conv3_fwd": conv3, "conv3_bwd": conv3,
            "reparam_fwd": rep, "reparam_bwd": rep}
    return {k: 2.0 * v * items for k, v in macs.items()}


# ----------------------------------------------------------------------------- #
# CPU arm: the UNMODIFIED reference (oracle/_ref) through its own public API on the host cores;
# the class-level port (oracle/ref_port.py) only if the copy did not travel
# ----------------------------------------------------------------------------- #
def cpu_reference(steps: int, warmup: int, rows: int = 64, pairs: int = 8, short: bool = False):
    from oracle import build_ref
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    wl = make_workload(rows, seed=123)
    out = dict(rows=rows, threads=threads)
    if build_ref.available():
        from oracle import ref_arm
        ref = build_ref.load()
        grouping = ref.prior_model.get_grouping_by_kl(wl["bits"])
        G = grouping[5]
        m = ref_arm.build_model(wl["cfg"], "cifar", rows, wl["A"], wl["up"], wl["p_loc"], wl["p_log_scale"], grouping)
        t_step = ref_arm.time_fit(m, wl["x"].contiguous(), wl["y"], steps, warmup)
        t_table, t_pair = ref_arm.time_rec(m, pairs)
        kind = "reference"
        if short:
            ws = make_workload(SHORT["rows"], seed=321)
            ms = ref_arm.build_model(ws["cfg"], "cifar", SHORT["rows"], ws["A"], ws["up"], ws["p_loc"], ws["p_log_scale"], grouping)
            d, wall, kl_bits = ref_arm.short_schedule(ms, ws["x"].contiguous(), ws["y"], SHORT["n_fit"], SHORT["n_finetune"])
            out["short"] = dict(rows=SHORT["rows"], fit_steps=SHORT["n_fit"], finetune_steps=SHORT["n_finetune"], rounds=G,
                                wall_s=wall, datapoints_per_s=SHORT["rows"] / wall, psnr_db=float(np.mean(d)),
                                bpp=float(ms.bpp), kl_bits_after_fit=float(np.mean(kl_bits)),
                                api="reference TestBNNmodel.optimize_posteriors + compress_posteriors, device='cpu'")
    else:
        from oracle.ref_port import OracleCompressor
        case = oracle_case(wl)
        G = case["lvl1"]["n_groups"]
        oc = OracleCompressor(case)
        for ep in range(warmup):
            oc.fit_step(ep, S)
        t0 = time.perf_counter()
        for ep in range(steps):
            oc.fit_step(warmup + ep, S)
        t_step = (time.perf_counter() - t0) / max(steps, 1)
        oc.gumbel()
        blocks = list(range(min(pairs, G)))
        t0 = time.perf_counter()
        for b in blocks:
            oc.table(b)
        t_table = (time.perf_counter() - t0) / len(blocks)
        t0 = time.perf_counter()
        for i, b in enumerate(blocks):
            oc.code_block(i % rows, b)
        t_pair = (time.perf_counter() - t0) / len(blocks)
        kind = "port"
    steps_full = schedule(G)
    total = steps_full * t_step + G * rows * t_pair + G * t_table
    out.update(value=rows / total, t_step=t_step, t_pair=t_pair, t_table=t_table, G=G, kind=kind,
               cand_per_s=N_CAND / t_pair,
               sample=f"{rows} cifar-shape rows, S=5: {steps} train() steps after {warmup} warm-ups + {min(pairs, G)} "
                      f"sample_group (row,block) pairs on {threads} threads ({'unmodified reference classes' if kind == 'reference' else 'CPU port'}), "
                      f"extrapolated to the {steps_full}-step/{G}-round schedule")
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference(max(args.steps, 1), max(args.warmup, 1), rows=64, short=not args.no_short)
    steps_full = schedule(r["G"])
    line = {"impl": "reference", "metric": "datapoints compressed/sec (CIFAR-10 32x32)", "value": r["value"],
            "unit": "datapoints/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["t_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 fit / f64 REC", "data": "synthetic",
            "config": {"workload": "cifar-shape 32x32, 64-row bounded sample of the 1024-row batch, S=5, G=%d (0.52 bpp)" % r["G"],
                       "schedule_steps": steps_full, "rec_rounds": r["G"]},
            "cpu_baseline": {"value": r["value"], "unit": "datapoints/s", "cores": r["threads"], "kind": r["kind"],
                             "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "datapoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "rec": {"candidates_per_s": r["cand_per_s"], "s_per_pair": r["t_pair"], "table_build_s_per_block": r["t_table"]},
            "gpu_launches": 0}
    if "short" in r:
        line["e2e_short"] = r["short"]
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- #
# prior training (north_star (c)): full-batch Adam step over the rank's rows with the weight gradients of the
# shared mappings (all-reduced over ranks), S = 1 -- one iteration of prior_model.py:229-262
# ----------------------------------------------------------------------------- #
def prior_training_ms(wl, dev, steps=10, warmup=3):
    from recombiner_b200.prior_model import PriorBNNmodel, LinearTransform, Upsample
    cfg, dims, rows = wl["cfg"], wl["dims"], wl["rows"]
    torch.manual_seed(7)
    lt = LinearTransform(dims).to(dev)
    up = Upsample(cfg["data_dim"], cfg["paddings"], cfg["layerwise_scale_factors"]).to(dev)
    with contextlib.redirect_stdout(io.StringIO()):
        m = PriorBNNmodel(in_dim=dims[0], hidden_dims=dims[1:-1], out_dim=dims[-1], train_size=rows,
                          data_dim=cfg["data_dim"], pixel_sizes=cfg["pixel_sizes"], upsample_factors=cfg["upsample_factors"],
                          latent_dim=cfg["latent_dim"], patch=cfg["patch"], patch_nums=cfg["patch_nums"],
                          hierarchical_patch_nums=cfg["hierarchical_patch_nums"], device=dev,
                          layer_scales=cfg["layerwise_scale_factors"], paddings=cfg["paddings"])
    grid = [p // u for p, u in zip(cfg["pixel_sizes"], cfg["upsample_factors"])]
    pri = (torch.zeros(wl["W"]), torch.full((wl["W"],), 0.02), torch.zeros(*grid, cfg["latent_dim"]),
           torch.full((*grid, cfg["latent_dim"]), 0.02))
    x, y = wl["x"].to(dev), wl["y"].to(dev)

    def run(n):
        m.train(n, 2e-4, x, y, *pri, None, None, None, None, lt, up, 1e-8, True, False)
    run(warmup)
    run(steps)              # a full-length untimed call: per-call setup (Adam state, allocator growth) is paid here
    ts = []
    for _ in range(3):      # median of three timed calls: this leg allocates per step and is sensitive to allocator state
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(steps)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / steps)
    return sorted(ts)[1]


# ----------------------------------------------------------------------------- #
# this repo's arm
# ----------------------------------------------------------------------------- #
def fp64_peak_tflops(dev):
    """DFMA peak of this device, measured live (rcb_ubench_dfma: 16 independent chains per thread)."""
    import ctypes as C
    from recombiner_b200 import _lib
    lib = _lib.load()
    scratch = torch.zeros(8, dtype=torch.float64, device=dev)
    flop = C.c_double(0.0)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.rcb_ubench_dfma(scratch.data_ptr(), 148 * 8, 2000, C.byref(flop), st), "rcb_ubench_dfma")
    best = 0.0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.rcb_ubench_dfma(scratch.data_ptr(), 148 * 8, 20000, C.byref(flop), st), "rcb_ubench_dfma")
        e1.record()
        torch.cuda.synchronize()
        best = max(best, flop.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def short_schedule_gpu(dev, rows, seed, rank=0):
    """optimize_posteriors + compress_posteriors through the public API with HOST inputs, wall clock."""
    from recombiner_b200.utils import batch_PSNR
    ws = make_workload(rows, seed=seed)
    m = build_model(ws, dev, row_offset=rank * rows)
    x, y = ws["x"][:1].clone().pin_memory().expand(rows, -1, -1), ws["y"].pin_memory()
    m._ensure_rec(N_CAND)                           # candidate tables: one-off per (prior, seed), like the CPU arm's cache
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        m.optimize_posteriors(x, y, n_epochs=SHORT["n_fit"], lr=2e-4, verbose=0)
        kl_bits = m.update_annealing_factors(False).sum(1) / np.log(2.)
        d = m.compress_posteriors(x, y, n_epochs_finetune=SHORT["n_finetune"], verbose=0, lr=2e-4, fine_tune_gap=1,
                                  compress_from_group_with_largest_kl=True)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    return dict(rows=rows, fit_steps=SHORT["n_fit"], finetune_steps=SHORT["n_finetune"], rounds=int(m.n_groups),
                wall_s=wall, datapoints_per_s=rows / wall, psnr_db=float(np.mean(d)), bpp=float(m.bpp),
                kl_bits_after_fit=float(np.mean(kl_bits)),
                api="TestBNNmodel.optimize_posteriors + compress_posteriors, host tensors in, distortion out")


def run_b200(args):
    import torch.distributed as dist
    from recombiner_b200 import _lib
    from recombiner_b200.engine import SectionTimer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    wl = make_workload(ROWS_PER_GPU, seed=1000 + rank)
    m = build_model(wl, dev, row_offset=rank * ROWS_PER_GPU)
    G = wl["G"]
    x_host = wl["x"][:1].clone().pin_memory().expand(ROWS_PER_GPU, -1, -1)     # Fourier inputs: one row, broadcast
    y_host = wl["y"].pin_memory()
    x, y = x_host[:1].to(dev).expand(ROWS_PER_GPU, -1, -1), wl["y"].to(dev)
    opt = torch.optim.Adam(m.parameters(), lr=2e-4)      # only its hyper-parameters are read (as in optimize_posteriors)
    cfg = m._adam_config(opt)
    m._ensure_rec(N_CAND)                                # candidate tables + Gumbel built once per prior

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, Wm = max(args.steps, 1), max(args.warmup, 3)
    step_no = [0]

    def fit_steps(n):
        for _ in range(n):
            m.fit_step(x, y, step_no[0], cfg, S)
            step_no[0] += 1

    # ---- device-resident timing (value)
    fit_steps(Wm)

    def warm_variants(xv, yv):
        """Untimed: a step with and a step without the beta update, twice each -- a configuration's step is captured
        into a CUDA graph the second time it is seen and replayed from then on."""
        for anneal in (True, False, True, False):
            m.fit_step(xv, yv, step_no[0], cfg, S, anneal=anneal)
            step_no[0] += 1
    warm_variants(x, y)
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    _lib.COUNTERS["launches"] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fit_steps(K)                                          # as train() runs them: replayed from the captured step
    e1.record()
    barrier()
    t_fit = e0.elapsed_time(e1) / K                       # ms per step
    launches = _lib.COUNTERS["launches"]
    # per-kernel times for the roofline: the same steps launched one by one with event brackets
    m.engine.timer = SectionTimer()
    fit_steps(K)
    sections = m.engine.timer.summary()
    m.engine.timer = None

    # ---- REC rounds (each codes one block of every row).  (a) the blocks the rows pick themselves (largest KL:
    # with this random-init workload nearly all rows pick the same few blocks, tables stay in L2); (b) worst case for
    # table traffic: the rows spread evenly over all G blocks (scored, not committed).
    n_rounds = min(max(2, min(K, 6)), G - 2)
    m.compress_round()                                    # warm-up round
    barrier()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _lib.COUNTERS["launches"] = 0
    r0.record()
    for _ in range(n_rounds):
        m.compress_round()
    r1.record()
    barrier()
    t_round = r0.elapsed_time(r1) / n_rounds
    launches += _lib.COUNTERS["launches"]
    spread = (torch.arange(ROWS_PER_GPU, device=dev, dtype=torch.int32) % G).contiguous()
    m.compress_round(blocks=spread, apply=False)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(n_rounds):
        m.compress_round(blocks=spread, apply=False)
    s1.record()
    barrier()
    t_round_spread = s0.elapsed_time(s1) / n_rounds
    f64_peak = fp64_peak_tflops(dev)

    # ---- end to end through the public call with HOST buffers: every step is one `train(x_host, y_host, n_epochs=1)`
    # (the loop main_compression drives, fed one step per call).  train() uploads that call's inputs from pinned memory
    # on a copy stream into one of two staging sets, so step k+1's upload travels under step k's kernels; the step's
    # loss terms come back to the host one step late, also on a side stream.
    sq_host = [torch.empty(ROWS_PER_GPU * S).pin_memory() for _ in range(2)]
    sq_stage = [torch.empty(ROWS_PER_GPU * S, device=dev) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream()
    step_done = [torch.cuda.Event() for _ in range(2)]
    back_done = [torch.cuda.Event() for _ in range(2)]
    loss_sum = 0.0

    def public_step():
        m.train(x_host, y_host, n_epochs=1, optimizer=opt, verbose=False, sample_size=S, start_epoch=step_no[0])
        step_no[0] += 1

    for _ in range(8):                                    # untimed: both staging sets see both step variants twice
        step_no[0] = (step_no[0] // 10 + 1) * 10          # ... an annealing step
        public_step(); public_step()
        step_no[0] += 1
        public_step(); public_step()                      # ... and plain ones
    step_no[0] = (step_no[0] // 10 + 1) * 10
    barrier()
    n_before = _lib.COUNTERS["launches"]
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for k in range(K):
        b = k & 1
        public_step()
        main_stream.wait_event(back_done[b])              # the staging copy of two steps ago has been read back
        sq_stage[b].copy_(m.last_loss_terms(S), non_blocking=True)
        step_done[b].record(main_stream)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(step_done[b])
            sq_host[b].copy_(sq_stage[b], non_blocking=True)
            back_done[b].record(copy_stream)
        if k > 0:                                         # the previous step's loss terms, on the host
            back_done[b ^ 1].synchronize()
            loss_sum += float(sq_host[b ^ 1][0])
    back_done[(K - 1) & 1].synchronize()
    loss_sum += float(sq_host[(K - 1) & 1][0])
    f1.record()
    barrier()
    clk = clocks.stop() if clocks else None            # sampled over the fit, REC and end-to-end regions
    t_e2e = f0.elapsed_time(f1) / K
    h2d = x_host[:1].numel() * 4 + y_host.numel() * 4
    d2h = sq_host[0].numel() * 4

    # ---- one short schedule run to completion through the public calls (wall clock, host inputs)
    short = None
    if not args.no_short:
        short = {"same_rows_as_cpu_arm": short_schedule_gpu(dev, SHORT["rows"], seed=321, rank=rank),
                 "full_batch": short_schedule_gpu(dev, ROWS_PER_GPU, seed=2000 + rank, rank=rank)}
        short["full_batch"]["wall_s"] = float(_max_over_ranks(short["full_batch"]["wall_s"], dev, world))
        short["full_batch"]["datapoints_per_s"] = ROWS_PER_GPU * world / short["full_batch"]["wall_s"]

    try:
        t_prior = prior_training_ms(wl, dev)
    except Exception as exc:                               # reported, never fatal for the headline metric
        t_prior = float("nan")
        if rank == 0:
            print("bench.py: prior-training leg failed: %r" % (exc,), file=sys.stderr)
    times = torch.tensor([t_fit, t_round, t_e2e, t_round_spread, t_prior if t_prior == t_prior else -1.0], device=dev,
                         dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_fit, t_round, t_e2e, t_round_spread, t_prior = (float(v) for v in times.cpu())
    if t_prior < 0:
        t_prior = float("nan")

    if rank == 0:
        pk = peaks()
        steps_full = schedule(G)
        n_total = ROWS_PER_GPU * world
        value = n_total / ((steps_full * t_fit + G * t_round) * 1e-3)
        e2e = n_total / ((steps_full * t_e2e + G * t_round) * 1e-3)
        flops = section_flops(ROWS_PER_GPU * S)
        timed = {k: v[1] for k, v in sections.items()}
        dom = max((k for k in timed if k in flops), key=lambda k: timed[k])
        achieved = flops[dom] / (timed[dom] * 1e-3) / 1e12
        step_flops = sum(flops[k] for k in timed if k in flops)
        cand_per_s = ROWS_PER_GPU * world * N_CAND / (t_round * 1e-3)
        # REC scoring: 2 DFMA per candidate-dimension and row (the quadratic form), f64 like the reference
        rec_flop = 2.0 * 2.0 * ROWS_PER_GPU * wl["P"] / G * N_CAND
        table_bytes = ROWS_PER_GPU * (wl["P"] / G) * N_CAND * 4
        cpu = cpu_reference(3, 1, rows=64) if not args.no_cpu_baseline else None

        def rec_roof(t_ms, traffic, note):
            return {"bound": "fp64", "kernel": "rec_encode_staged_kernel", "achieved": rec_flop / (t_ms * 1e-3) / 1e12,
                    "peak": f64_peak, "unit": "TFLOP/s", "frac": rec_flop / (t_ms * 1e-3) / 1e12 / f64_peak,
                    "peak_source": "measured live: rcb_ubench_dfma, 16 independent DFMA chains per thread",
                    "table_GBps": table_bytes / (t_ms * 1e-3) / 1e9, "traffic": traffic, "note": note}
        line = {
            "metric": "datapoints compressed/sec (CIFAR-10 32x32)", "value": value, "unit": "datapoints/s",
            "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": t_fit, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": (("fp16 / tf32 operands" if m.engine.half_hw else "tf32 + fp16 operands")
                      + " (tcgen05, fp32 accumulate in TMEM; fp32 state and gradients) fit / f64 REC"
                      if m.engine.tc else "f32 fit (SIMT FFMA) / f64 REC"),
            "data": "synthetic",
            "config": {"workload": "cifar-shape 32x32, %d rows/GPU, S=5, G=%d blocks x 16 bit (0.52 bpp), "
                                   "random-init prior" % (ROWS_PER_GPU, G),
                       "schedule_steps": steps_full, "rec_rounds": G, "rec_ms_per_round": t_round,
                       "precision": m.engine.precision,
                       "timed": "K fit steps + %d REC rounds, CUDA events, max over ranks" % n_rounds,
                       "l2": "per-step working set ~2 GB > 126 MB L2 (no flush needed)",
                       "step_algorithmic_tflop": step_flops / 1e12,
                       "step_tflops": step_flops / (t_fit * 1e-3) / 1e12},
            "e2e": {"value": e2e, "unit": "datapoints/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": t_e2e,
                    "api": "TestBNNmodel.train(x_host, y_host, n_epochs=1, optimizer) per step; loss terms read back per step"},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": pk["tf"], "unit": "TFLOP/s",
                         "frac": achieved / pk["tf"], "traffic": TRAFFIC.get(dom), "peak_source": pk["src"] + " bf16 sustained",
                         "note": "MLP operands are fp16 (kind::f16, fp32 accumulation in TMEM): the kernel is bound by its "
                                 "CUDA-core epilogues (96 sin + 96 cos per pixel and tile, issue slots), not by the tensor "
                                 "pipe; executed FLOPs; per-launch ms: "
                                 + ", ".join(f"{k}={v:.3f}" for k, v in sorted(timed.items(), key=lambda kv: -kv[1]))},
            # the upsampler kernels are HBM-bound: algorithmic bytes per launch (tensors read + written once, in the
            # precision they are stored in) over the event-timed launch duration, against the measured copy bandwidth
            "hbm_kernels": {k: {"ms": timed[k], "algorithmic_MB": b / 1e6, "achieved_GBps": b / (timed[k] * 1e-3) / 1e9,
                                "frac": b / (timed[k] * 1e-3) / 1e9 / pk["hbm"]}
                            for k, b in conv_bytes(m.engine, ROWS_PER_GPU * S).items() if k in timed},
            # REC round: f64 quadratic form per (row, candidate, dimension); bound by the FP64 pipe once the tables
            # are staged through shared memory (they are shared by the rows of a run and mostly L2-resident)
            "rec": {"candidates_per_s": cand_per_s, "ms_per_round": t_round, "pairs_per_round": ROWS_PER_GPU * world,
                    "roofline": rec_roof(t_round, REC_TRAFFIC["picked"], "blocks picked by the rows themselves (largest KL): with the "
                                         "random-init workload they coincide, so the candidate tables are L2-resident"),
                    "spread_over_all_blocks": dict(rec_roof(t_round_spread, REC_TRAFFIC["spread"], "rows spread evenly over all G blocks: "
                                                            "every table (0.99 GB in all) is streamed from HBM at least once"),
                                                   ms_per_round=t_round_spread,
                                                   candidates_per_s=ROWS_PER_GPU * world * N_CAND / (t_round_spread * 1e-3))},
            "prior_training": {"ms_per_step": t_prior, "rows_per_gpu": ROWS_PER_GPU,
                               "rows_per_s": (ROWS_PER_GPU * world / (t_prior * 1e-3)) if t_prior == t_prior else None,
                               "note": "full-batch Adam step, S=1, weight gradients of A and the upsampler included; median of 3 timed calls"
                                       + (", all-reduced over %d ranks (NCCL)" % world if world > 1 else "")},
            "clocks": clk,
        }
        if short:
            line["e2e_short"] = short
        if cpu:
            line["cpu_baseline"] = {"value": cpu["value"], "unit": "datapoints/s", "cores": cpu["threads"], "kind": cpu["kind"],
                                    "sample": cpu["sample"], "rec_candidates_per_s": cpu["cand_per_s"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- #
# other modality shapes (BASELINE.json configs 3-5): one fit step, driver-visible numbers
# ----------------------------------------------------------------------------- #
MODALITY_DATA = {"kodak": 1, "audio": 4, "video": 1, "protein": 1000}     # data per GPU (kodak / video: the reference's 1-datum batch)


def modality_flops(eng, rows):
    """Executed FLOPs of one fit step (S samples, forward + data gradients; polyphase convolutions)."""
    items, citems = rows * S, (rows // eng.R) * S
    conv = 0
    for i, g in enumerate(eng.geoms):
        out_px = g.d * g.fz * g.h * g.fy * g.w * g.fx
        taps = (1 if g.kz == 1 else 2) * (1 if g.ky == 1 else 2) * (1 if g.kx == 1 else 2)
        if i == 0 and eng.dense1:
            conv += (g.h * g.w * g.ic) * (out_px * g.oc)
        else:
            conv += out_px * g.oc * taps * g.ic
    rep = sum(c * c for c in eng.counts)
    d0 = eng.dims[0]
    mlp_fwd = eng.pix * (d0 * 32 + 2 * 32 * 32 + 32 * eng.out)
    mlp_bwd = mlp_fwd + eng.pix * (32 * eng.out + 2 * 32 * 32 + 16 * 32)
    return 2.0 * (2 * conv * citems + 2 * rep * items + (mlp_fwd + mlp_bwd) * items)


def run_modality(args):
    from recombiner_b200 import _lib
    from recombiner_b200.engine import SectionTimer
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    if int(os.environ.get("RANK", "0")) != 0:
        return
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    name = args.workload
    from recombiner_b200.config import configs
    cfg = configs[name]
    R = int(np.prod(cfg["patch_nums"])) if cfg["patch"] else 1
    rows = MODALITY_DATA[name] * R
    wl = make_workload(rows, seed=77, dataset=name)
    m = build_model(wl, dev)
    x, y = wl["x"][:1].to(dev).expand(rows, -1, -1), wl["y"].to(dev)
    opt = torch.optim.Adam(m.parameters(), lr=2e-4)
    cfg_adam = m._adam_config(opt)
    K, Wm = max(args.steps, 1), max(args.warmup, 3)
    step = [0]

    def steps(n, anneal=None):
        for _ in range(n):
            m.fit_step(x, y, step[0], cfg_adam, S, anneal=anneal)
            step[0] += 1
    steps(Wm)
    for an in (True, False, True, False):
        steps(1, an)
    torch.cuda.synchronize()
    clocks = ClockSampler(dev.index or 0)
    _lib.COUNTERS["launches"] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    steps(K)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / K
    launches = _lib.COUNTERS["launches"]
    m.engine.timer = SectionTimer()
    steps(K)
    timed = {k: v[1] for k, v in m.engine.timer.summary().items()}
    m.engine.timer = None
    clk = clocks.stop()
    pk = peaks()
    fl = modality_flops(m.engine, rows)
    unit = {"kodak": "images", "audio": "clips", "video": "clips", "protein": "chains"}[name]
    line = {"metric": "fit steps/s (%s shape)" % name, "value": 1e3 / t, "unit": "steps/s", "n_gpus": 1, "steps": K, "warmup": Wm,
            "ms_per_step": t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32 + fp16 operands (tcgen05) / fp32 state" if m.engine.tc else "f32", "data": "synthetic",
            "config": {"workload": "%s shape: %d %s = %d rows, S=5, one fit step (test_model.py:622-635)" % (name, MODALITY_DATA[name], unit, rows),
                       "mlp_kernel": "tcgen05" if (m.engine.tc_mlp and m.engine.n_f in (16, 18)) else "simt",
                       "step_algorithmic_tflop": fl / 1e12},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "whole fit step", "achieved": fl / (t * 1e-3) / 1e12, "peak": pk["tf"],
                         "unit": "TFLOP/s", "frac": fl / (t * 1e-3) / 1e12 / pk["tf"], "traffic": None,
                         "peak_source": pk["src"] + " bf16 sustained",
                         "note": "executed FLOPs of the step over its replayed duration; per-launch ms: "
                                 + ", ".join(f"{k}={v:.3f}" for k, v in sorted(timed.items(), key=lambda kv: -kv[1]))},
            "clocks": clk}
    print(json.dumps(line), flush=True)


def _max_over_ranks(v, dev, world):
    import torch.distributed as dist
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def _quiet_stdout():
    """Libraries (NCCL prints its version banner) write to file descriptor 1; the contract is ONE JSON line on
    stdout.  Point fd 1 at stderr for the run and keep the real stdout for the final line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-short", action="store_true", help="skip the e2e_short leg (a short schedule run to completion)")
    ap.add_argument("--workload", default="cifar", choices=["cifar", "kodak", "audio", "video", "protein"],
                    help="cifar = the headline bench; the others time one fit step of that modality's shape")
    args = ap.parse_args()
    real_stdout = _quiet_stdout()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        if args.impl == "reference":
            run_reference(args)
        elif args.workload != "cifar":
            run_modality(args)
        else:
            run_b200(args)
    lines = [ln for ln in buf.getvalue().splitlines() if ln.startswith("{")]
    other = [ln for ln in buf.getvalue().splitlines() if not ln.startswith("{")]
    if other:
        print("\n".join(other), file=sys.stderr)
    if lines:
        real_stdout.write(lines[-1] + "\n")
        real_stdout.flush()


if __name__ == "__main__":
    main()
