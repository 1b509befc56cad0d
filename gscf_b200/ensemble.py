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


class NcclGather:
    """The C-ABI collective (gscf_b200_comm_* / gscf_b200_ensemble_gather, include/gscf_b200.h section 3b) driven from
    a torch.distributed job: rank 0 creates the NCCL unique id, torch.distributed only carries those 128 bytes to the
    other ranks; the gather itself is the library's own ncclAllGather on the engine's stream."""

    def __init__(self, group=None):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import lib as L

        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        buf = C.create_string_buffer(128)
        ok, why = 1, ""
        if self.rank == 0:
            st = L.lib.gscf_b200_comm_unique_id(buf)
            if st != L.OK:
                ok, why = 0, L.last_error()
        if self.world > 1:
            # every rank takes part in this broadcast whatever happened on rank 0 (byte 0 = "rank 0 has an id"), so a
            # missing NCCL library surfaces as the same exception everywhere instead of a hang
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
            t = torch.frombuffer(bytearray(bytes([ok]) + buf.raw), dtype=torch.uint8).to(dev)
            dist.broadcast(t, src=0, group=group)
            raw = bytes(t.cpu().numpy().tobytes())
            ok = raw[0]
            buf = C.create_string_buffer(raw[1:], 128)
        if not ok:
            raise RuntimeError("rank 0 could not create an NCCL unique id: " + (why or "see rank 0"))
        self.handle = C.c_void_p()
        L.check(L.lib.gscf_b200_comm_create(buf, self.world, self.rank, C.byref(self.handle)))

    def gather(self, engine, n_problems: int) -> np.ndarray:
        """Records of all problems in problem order (P x 5, RECORD_FIELDS); `engine` holds this rank's shard."""
        import ctypes as C

        from . import lib as L

        counts = (C.c_int * self.world)(*[hi - lo for lo, hi in (shard_range(n_problems, r, self.world)
                                                                    for r in range(self.world))])
        out = (L.Record * n_problems)()
        L.check(L.lib.gscf_b200_ensemble_gather(engine.handle, self.handle, counts, out))
        return np.array([[r.energy_total, r.error_norm, r.n_iter, r.converged, r.status] for r in out], dtype=float
                        ).reshape(n_problems, len(RECORD_FIELDS))

    def close(self):
        from . import lib as L

        if self.handle:
            L.lib.gscf_b200_comm_destroy(self.handle)
            self.handle = None


def _solve_shard_on_gpu(n_bas, n_alpha, n_beta, n_aux, seeds, method, params):
    from .solvers import DeviceDFFock, solve_device_problem

    fock = DeviceDFFock(n_bas, n_alpha, n_beta, n_aux=n_aux, seeds=seeds)
    eng = None
    try:
        eng, st = solve_device_problem(fock, method, params)
        e1, e2 = eng.energies()
        return pack_records(e1 + e2, eng.error_norm(), eng.n_iter(), eng.converged(), eng.status())
    finally:
        if eng is not None:
            eng.close()
        fock.close()


def solve_ensemble(n_bas, n_alpha, n_beta=None, n_aux=None, n_problems=1, seeds: Optional[List[int]] = None,
                   method="diis", params=None, group=None, local_solver=None):
    """Solve this rank's shard of a synthetic DF ensemble on its GPU and gather the records of all problems (problem
    order).  Every rank reaches the collective exactly once: a rank whose shard is empty (more ranks than problems)
    contributes zero records, and a rank whose local solve raises contributes records with status -1 and re-raises
    after the gather, so the other ranks never hang in it.  `local_solver(seeds) -> records` replaces the GPU solve
    (host-logic tests)."""
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    lo, hi = shard_range(n_problems, rank, world)
    all_seeds = list(range(n_problems)) if seeds is None else list(seeds)
    failure = None
    if hi == lo:
        local = np.zeros((0, len(RECORD_FIELDS)))
    else:
        try:
            if local_solver is not None:
                local = np.asarray(local_solver(all_seeds[lo:hi]), dtype=float)
            else:
                local = _solve_shard_on_gpu(n_bas, n_alpha, n_beta, n_aux, all_seeds[lo:hi], method, params)
        except Exception as exc:  # noqa: BLE001 - the collective below must still be reached
            failure = exc
            local = np.zeros((hi - lo, len(RECORD_FIELDS)))
            local[:, 1] = np.nan
            local[:, 4] = -1.0
    records = gather_records(local, n_problems, group)
    if failure is not None:
        raise failure
    return records
