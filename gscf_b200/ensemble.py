"""Multi-GPU ensembles: one process per GPU, problems sharded by index, results gathered.

Only the batched ensemble of independent SCF problems is distributed: rank r of R
owns the contiguous block [r*P/R, (r+1)*P/R) of problem indices for the whole run, there is no
data-path collective, and the single collective is the gather of the per-problem records
{E_total, ||e||, n_iter, status} at the end (NCCL over NVLink on the GPU box, gloo in the CPU
tests).  A single large problem is never split across GPUs.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

RECORD_FIELDS = ("energy", "error_norm", "n_iter", "converged", "status")


def shard_range(n_problems: int, rank: int, world: int) -> Tuple[int, int]:
    """Static contiguous block partition; the first (P mod R) ranks get one extra problem."""
    base, extra = divmod(n_problems, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def pack_records(energy, error_norm, n_iter, converged, status) -> np.ndarray:
    rec = np.stack([np.asarray(energy, float), np.asarray(error_norm, float), np.asarray(n_iter, float),
                    np.asarray(converged, float), np.asarray(status, float)], axis=1)
    return np.ascontiguousarray(rec)


def gather_records(local: np.ndarray, n_problems: int, group=None, device=None) -> np.ndarray:
    """All-gather the per-problem records of every rank into problem order (P x 5)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_problems, r, world) for r in range(world)]
    maxlen = max(hi - lo for lo, hi in sizes)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
    buf = torch.zeros((maxlen, len(RECORD_FIELDS)), dtype=torch.float64, device=device)
    lo, hi = sizes[rank]
    assert local.shape == (hi - lo, len(RECORD_FIELDS)), (local.shape, hi - lo)
    buf[: hi - lo] = torch.from_numpy(local).to(device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    parts = [o[: h - l].cpu().numpy() for o, (l, h) in zip(out, sizes)]
    return np.concatenate(parts, axis=0)


def summarise(records: np.ndarray) -> Dict[str, float]:
    return {
        "n_problems": int(records.shape[0]),
        "n_converged": int(records[:, 3].sum()),
        "iterations_total": int(records[:, 2].sum()),
        "iterations_max": int(records[:, 2].max()) if len(records) else 0,
        "iterations_mean": float(records[:, 2].mean()) if len(records) else 0.0,
    }


def solve_ensemble(n_bas, n_alpha, n_beta=None, n_aux=None, n_problems=1, seeds: Optional[List[int]] = None,
                   method="diis", params=None, group=None):
    """Solve this rank's shard of a synthetic DF ensemble on its GPU and gather the records.

    Returns (records of all problems in problem order, local Engine closed)."""
    import torch.distributed as dist

    from . import lib as L
    from .solvers import DeviceDFFock, solve_device_problem

    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    lo, hi = shard_range(n_problems, rank, world)
    all_seeds = list(range(n_problems)) if seeds is None else list(seeds)
    fock = DeviceDFFock(n_bas, n_alpha, n_beta, n_aux=n_aux, seeds=all_seeds[lo:hi])
    eng, st = solve_device_problem(fock, method, params)
    try:
        e1, e2 = eng.energies()
        local = pack_records(e1 + e2, eng.error_norm(), eng.n_iter(), eng.converged(), eng.status())
    finally:
        eng.close()
        fock.close()
    return gather_records(local, n_problems, group)
