"""Data-parallel plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in
the CPU tests).  Compression shards rows with no data-path collective; prior training
all-reduces shared-mapping gradients per step and f64 sufficient statistics per EM
iteration."""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def shard_rows(n_rows: int, world_size: int, rank: int, unit: int = 1):
    """Contiguous [start, end) block of rows for `rank`; `unit` rows stay together (patch
    modalities shard by whole datum: 96 / 60 / 64 patches).  Remainder units go to the
    first ranks."""
    if n_rows % unit:
        raise ValueError(f"{n_rows} rows is not a multiple of the shard unit {unit}")
    units = n_rows // unit
    base, extra = divmod(units, world_size)
    start = rank * base + min(rank, extra)
    count = base + (1 if rank < extra else 0)
    return start * unit, (start + count) * unit


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks (no-op for a single process)."""
    if world()[0] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def prior_from_stats(stats: torch.Tensor, n_total: int):
    """Host restatement of rcb_prior_from_stats for checks on any device:
    stats = [sum mu, sum mu^2, sum sigma^2] (f64, already all-reduced)."""
    P = stats.numel() // 3
    s0, s1, s2 = stats[:P], stats[P:2 * P], stats[2 * P:]
    mean = s0 / n_total
    var_mu = (s1 - n_total * mean * mean) / (n_total - 1)
    return mean, torch.sqrt(torch.clamp(s2 / n_total + var_mu, min=0.0))


def gather_rows(local: torch.Tensor) -> torch.Tensor:
    """Concatenate per-rank row blocks on every rank (result collection at the end of a
    sharded compression: indices / distortions, a few KB)."""
    w, _ = world()
    if w == 1:
        return local
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(w)]
    dist.all_gather(sizes, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device))
    mx = int(max(int(s.item()) for s in sizes))
    pad = torch.zeros(mx, *local.shape[1:], dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(w)]
    dist.all_gather(out, pad)
    return torch.cat([o[: int(s.item())] for o, s in zip(out, sizes)], 0)
