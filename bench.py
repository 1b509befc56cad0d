#!/usr/bin/env python
"""bench.py - SCF-iteration throughput of the gscf hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm   (libgscf_b200.so on the GPU)
  python bench.py --impl reference --steps K --warmup W    reference arm: the CPU restatement of the
                                                           reference algorithm (oracle/, LAPACK) on the
                                                           host cores - the reference itself cannot be
                                                           built here (un-vendored lazyten/krims).

A *step* is one complete SCF solve of one batch of synthetic input: PulayDiisScf to convergence from
the core guess the caller brings (computed once, outside the timed region, like the reference's
solve(fock, S) whose caller builds the initial matrix from a guess).  Default workload (N = 1): BASELINE configs[1] - synthetic closed-shell DF problem
n_bf = 1024, n_occ = 128, DIIS depth 8, fp64.  With N > 1 every rank solves an independent replica
(a single large problem stays on one GPU, north_star); the only collective is the NCCL gather of
per-problem energies / iteration counts / convergence flags.

metric  scf_iterations_per_s = SCF iterations / time in the SCF path, Fock builder excluded
        (SURVEY 8d).  `value` uses CUDA-event device time (orthogonaliser set-up + solve loop);
        `e2e` is the same metric with HOST input buffers: wall clock around
        set_overlap(host S) + set_guess(host C0) + solve + read-back of C, eps, energies into pinned
        host buffers, minus the event-timed Fock-builder time (loop and guess).
        Ensemble workloads (c4, c5) are sharded over the ranks (strong scaling); their CPU baseline
        runs one member per host core with single-threaded BLAS.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_bf, n_alpha, n_beta, n_aux, diis depth, problems per GPU)
    "c2": dict(n=1024, na=128, nb=128, aux=64, m=8, P=1, desc="configs[1]: n_bf=1024 n_occ=128 DIIS 8"),
    "c3": dict(n=4096, na=512, nb=512, aux=64, m=10, P=1, desc="configs[2]: n_bf=4096 n_occ=512 DIIS 10 (n_aux=64)"),
    # configs[2] as written: "DIIS 10 then TruncatedOptDampScf" - DIIS down to 1e-3, hand-over, TODA to 5e-7
    "c3t": dict(n=4096, na=512, nb=512, aux=64, m=10, P=1, chain=1e-3,
                desc="configs[2]: n_bf=4096 n_occ=512 DIIS 10 (to 1e-3) then TruncatedOptDampScf (n_aux=64)"),
    "c4": dict(n=128, na=16, nb=16, aux=8, m=4, P=4096, desc="configs[3]: 4096 x n_bf=128 ensemble, DIIS 4"),
    "c5": dict(n=512, na=64, nb=60, aux=32, m=4, P=256, desc="configs[4]: 256 x UHF n_bf=512 ensemble, DIIS 4"),
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=self.tmp,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.strip().split(",") for r in open(self.tmp.name) if r.strip()]
        os.unlink(self.tmp.name)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU side: oracle timed on the host cores (cpu_baseline leg and --impl reference)
# ---------------------------------------------------------------------------------------------
def oracle_time_solve(w, seed, max_iter=None):
    """One DIIS solve of the oracle; returns (n_iter, scf_path_seconds, fock_seconds, energy)."""
    from oracle import ScfFailed, pulay_diis_scf
    from oracle.synthetic import SyntheticDFProblem

    prob = SyntheticDFProblem(w["n"], w["na"], w["nb"], n_aux=w["aux"], seed=seed)
    fock_t = [0.0]
    inner = prob.fock

    def timed_fock(C):
        t0 = time.perf_counter()
        out = inner(C)
        fock_t[0] += time.perf_counter() - t0
        return out

    prob.fock = timed_fock
    C0 = prob.core_guess()
    C0 = np.stack([C0, C0]) if prob.unrestricted else C0[None]
    F0, e1, e2 = inner(C0)
    fock_t[0] = 0.0
    t0 = time.perf_counter()
    n_iter, energy = 0, float("nan")
    try:
        res = pulay_diis_scf(prob, F0, e1, e2, n_prev_steps=w["m"], max_iter=max_iter or 100)
        n_iter, energy = res.n_iter, res.energy_total
    except ScfFailed:
        n_iter = max_iter or 100  # bounded sample: the iterations were executed and timed
    total = time.perf_counter() - t0
    return n_iter, total - fock_t[0], fock_t[0], energy


def _pool_worker(job):
    """One ensemble member on one core (BLAS threads = 1); returns (n_iter, scf_path_seconds)."""
    try:
        import scipy.linalg  # noqa: F401  (load its OpenBLAS before limiting the pools)
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=1)
    except Exception:  # noqa: BLE001
        pass
    w, seed, max_iter = job
    n_it, t_scf, _, _ = oracle_time_solve(w, seed=seed, max_iter=max_iter)
    return n_it, t_scf


def oracle_time_ensemble(w, seeds, max_iter=None):
    """CPU baseline of an ensemble workload: one problem per core, BLAS threads = 1, all cores busy.
    Returns (iterations/s aggregated over the cores, cores, members solved)."""
    import multiprocessing as mp

    cores = os.cpu_count() or 1
    ctx = mp.get_context("spawn")  # the parent may hold a CUDA context: never fork it
    saved = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS")}
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = "1"  # inherited by the workers
    try:
        with ctx.Pool(cores) as pool:
            res = pool.map(_pool_worker, [(dict(w), int(s), max_iter) for s in seeds], chunksize=1)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    its = sum(r[0] for r in res)
    busy = sum(r[1] for r in res)          # core-seconds spent in the SCF path while every core was loaded
    return its / (busy / cores), cores, len(res), its


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    # torchrun exports OMP_NUM_THREADS=1 to its workers: give LAPACK / BLAS every host core back (the arm is "the reference's
    # CPU path with all the host threads it can use"; the other ranks have already returned)
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=cores)
    except Exception:  # noqa: BLE001
        pass
    if w["P"] > 1:
        # ensembles: one member per core; every step solves 2 members per core completely
        vals, times, iters = [], [], []
        for s in range(args.warmup + args.steps):
            v, cores, nmem, its = oracle_time_ensemble(w, range(2 * cores))
            if s >= args.warmup:
                vals.append(v)
                iters.append(its)
                times.append(its / v)  # SCF-path wall time of the step with every core busy
        value = sum(iters) / sum(times)
        sample = f"{2 * cores} members of {w['desc']} per step, one per core (BLAS threads = 1), full DIIS solves"
    else:
        sample_iters = 3 if w["n"] >= 1024 else 6
        sample = f"{sample_iters} DIIS iterations of one {w['desc']} problem per step (BLAS threads = all cores)"
        times, iters = [], []
        for s in range(args.warmup + args.steps):
            n_it, t_scf, t_fock, _ = oracle_time_solve(w, seed=0, max_iter=sample_iters)
            if s >= args.warmup:
                times.append(t_scf)
                iters.append(n_it)
        value = sum(iters) / sum(times)
    line = {
        "impl": "reference", "metric": "scf_iterations_per_s", "value": value, "unit": "iterations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["desc"], "impl": "oracle/ (numpy + scipy LAPACK dsygvd/dgesv/dgemm): the reference's "
                   "lazyten/armadillo path cannot be built (un-vendored dependencies)", "fock_builder": "excluded"},
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def ncu_traffic(n):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the committed
    `ncu --set full` capture of the same kernel at this order (round 2: profiles/r02_ncu_full_trd_{rega,blocked}.json,
    else round 1: profiles/r01_ncu_full_trd_fused.json), else None."""
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles")
    for name in ("r02_ncu_full_trd_rega.json", "r02_ncu_full_trd_blocked.json", "r01_ncu_full_trd_fused.json"):
        try:
            with open(os.path.join(here, name)) as f:
                rec = json.load(f)
        except (OSError, ValueError):
            continue
        got = rec.get(str(n), {}).get("dram_bytes")
        if got is not None:
            return got
    return None


class Job:
    """Process-wide context of our arm: device, ranks, the optional process group."""

    def __init__(self):
        import torch

        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (gscf_b200 has no CPU fallback)")
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        self.flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
        self.capi_gather = None
        self.gather_kind = "single rank"
        if self.world > 1:
            # the path's only collective through the library's own C ABI (gscf_b200_ensemble_gather: ncclAllGather on the
            # engine's stream); torch.distributed only carries the 128-byte NCCL id.  Falls back to torch's all_gather.
            try:
                from gscf_b200 import ensemble

                self.capi_gather = ensemble.NcclGather()
                self.gather_kind = "gscf_b200_ensemble_gather (C ABI, ncclAllGather)"
            except Exception as exc:  # noqa: BLE001
                self.capi_gather = None
                self.gather_kind = f"torch.distributed all_gather (C-ABI communicator unavailable: {exc})"
            ok = torch.tensor([1.0 if self.capi_gather is not None else 0.0], device="cuda")
            self.dist.all_reduce(ok, op=self.dist.ReduceOp.MIN)
            if float(ok.cpu()) < 1.0 and self.capi_gather is not None:
                self.capi_gather.close()
                self.capi_gather = None
                self.gather_kind = "torch.distributed all_gather (C-ABI communicator unavailable on some rank)"

    def close(self):
        if self.capi_gather is not None:
            self.capi_gather.close()
        if self.dist is not None:
            self.dist.destroy_process_group()


def measure(job, w, steps, warmup, cpu_baseline=True, sample_clocks=False, problems=0):
    """Time `steps` solves of workload `w` (after `warmup` untimed ones) on this job's GPUs.  Returns the result record
    on rank 0 (None elsewhere).  Single problems run as one independent replica per rank, ensembles are sharded by problem
    index (gscf_b200.ensemble.shard_range: no member dropped)."""
    import gscf_b200 as G
    from gscf_b200 import ensemble as ENS
    from gscf_b200 import lib as L

    torch, dist, world, rank = job.torch, job.dist, job.world, job.rank
    n, na, nb, aux, m = w["n"], w["na"], w["nb"], w["aux"], w["m"]
    is_ensemble = w["P"] > 1
    P_total = problems or w["P"]
    if is_ensemble:
        lo, hi = ENS.shard_range(P_total, rank, world)
        seeds = list(range(lo, hi))
    else:
        lo, hi = rank, rank + 1
        seeds = [rank]
    P_local = len(seeds)
    assert P_local >= 1, "more ranks than ensemble members"
    unrestricted = na != nb
    nblk = 2 if unrestricted else 1

    fock = G.DeviceDFFock(n, na, nb, n_aux=aux, seeds=seeds)
    ld = fock.ld
    # HOST copies of the inputs of the path (S and the core Hamiltonian), pinned
    S_host = torch.empty((P_local, n, n), dtype=torch.float64).pin_memory()
    H_host = torch.empty((P_local, n, n), dtype=torch.float64).pin_memory()
    for p in range(P_local):
        L.check(L.lib.gscf_b200_copy_matrix(S_host[p].data_ptr(), n, L.HOST, fock.overlap_ptr() + p * ld * n * 8, ld,
                                            L.DEVICE, n, n))
        L.check(L.lib.gscf_b200_copy_matrix(H_host[p].data_ptr(), n, L.HOST, fock.core_ptr() + p * ld * n * 8, ld,
                                            L.DEVICE, n, n))
    flush = job.flush

    params = L.Params()
    L.lib.gscf_b200_default_params(ctypes.byref(params))
    params.diis_n_prev_steps = m
    if w.get("n_eigenpairs"):
        params.n_eigenpairs = int(w["n_eigenpairs"])  # ScfBaseKeys::n_eigenpairs < n: partial-spectrum eigensolves
    fn_ptr, ctx = fock.device_callback()

    eng = G.Engine(n, na, nb, n_blocks=nblk, n_problems=P_local, max_history=max(m, 2))
    eng.set_fock_device(fn_ptr, ctx)
    eng.set_phase_timing(True)

    # pinned HOST result buffers: the C ABI getters copy straight into them (what a C++ host would do)
    C_host = torch.empty((P_local * nblk, n, n), dtype=torch.float64).pin_memory()
    w_host = torch.empty((P_local * nblk, n), dtype=torch.float64).pin_memory()
    e_host = torch.empty((2, P_local), dtype=torch.float64).pin_memory()
    _D = ctypes.POINTER(ctypes.c_double)

    def _p(t):
        return ctypes.cast(t.data_ptr(), _D)

    # the guess the caller brings (the reference's solve(fock, S) starts from a problem matrix built from a guess:
    # examples/hartree_fock/main.cc:57-76): core-guess coefficients, computed once outside the timed region, on the host
    G_host = torch.empty((P_local * nblk, n, n), dtype=torch.float64).pin_memory()
    L.check(L.lib.gscf_b200_set_overlap(eng.handle, S_host.data_ptr(), n, L.HOST))
    L.check(L.lib.gscf_b200_set_guess_from_matrix(eng.handle, H_host.data_ptr(), n, L.HOST))
    L.check(L.lib.gscf_b200_get_coeff(eng.handle, _p(G_host), n))
    torch.cuda.synchronize()

    def one_step():
        """One solve through the public C ABI with HOST inputs and HOST outputs; returns per-step measurements."""
        flush.zero_()  # L2 flush between timed iterations
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        L.check(L.lib.gscf_b200_set_overlap(eng.handle, S_host.data_ptr(), n, L.HOST))
        L.check(L.lib.gscf_b200_set_guess(eng.handle, G_host.data_ptr(), n, L.HOST, None))
        if w.get("chain"):
            # DIIS leg to a loose threshold, then TODA from that state (same engine: C, F, energies stay on the device -
            # the hand-over of ScfBase::solve_with_guess without the host round trip)
            p1 = L.Params()
            ctypes.memmove(ctypes.byref(p1), ctypes.byref(params), ctypes.sizeof(L.Params))
            p1.max_error_norm = float(w["chain"])
            st = eng.solve("diis", p1)
            its1 = eng.n_iter().copy()
            scf1, fock1 = eng.timings()
            if st == L.OK:
                st = eng.solve("toda", params)
        else:
            st = eng.solve("diis", params)
        L.check(L.lib.gscf_b200_get_energies(eng.handle, _p(e_host[0]), _p(e_host[1])))
        L.check(L.lib.gscf_b200_get_orben(eng.handle, _p(w_host)))
        L.check(L.lib.gscf_b200_get_coeff(eng.handle, _p(C_host), n))
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        its = eng.n_iter()
        e1, e2 = e_host[0].numpy().copy(), e_host[1].numpy().copy()
        scf_ms, fock_ms = eng.timings()
        if w.get("chain"):
            its = its + its1
            scf_ms, fock_ms = scf_ms + scf1, fock_ms + fock1
        ph = eng.phase_timings()
        fock_ms += ph["guess_fock"]  # F(C_guess): the plugin's time, like every other Fock build (the reference arm
        #                              builds F0 outside its timed region too)
        return dict(status=st, iters=its.copy(), e=(e1 + e2).copy(), conv=eng.converged().copy(), wall=wall,
                    scf_ms=scf_ms, fock_ms=fock_ms, overlap_ms=ph["overlap"], ph=ph,
                    d2h=C_host.numel() * 8 + w_host.numel() * 8 + e_host.numel() * 8)

    for _ in range(warmup):
        one_step()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(job.local_rank) if (sample_clocks and rank == 0) else None
    if sampler:
        sampler.start()
    launches0 = L.lib.gscf_b200_launch_count()
    t_begin = time.perf_counter()
    done = [one_step() for _ in range(steps)]
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t_total = time.perf_counter() - t_begin
    launches = L.lib.gscf_b200_launch_count() - launches0
    clocks = sampler.stop() if sampler else None

    iters_local = float(sum(int(s["iters"].sum()) for s in done))
    conv_local = float(sum(int(s["conv"].sum()) for s in done))
    dev_s = sum(s["scf_ms"] + s["overlap_ms"] for s in done) / 1e3       # device time in the SCF path
    e2e_s = sum(s["wall"] - s["fock_ms"] / 1e3 for s in done)            # host-buffer wall time, Fock excluded
    fock_s = sum(s["fock_ms"] for s in done) / 1e3
    last = done[-1]
    red = torch.tensor([dev_s, e2e_s, t_total, fock_s], dtype=torch.float64, device="cuda")
    tot = torch.tensor([iters_local, conv_local, float(launches)], dtype=torch.float64, device="cuda")
    per_rank = torch.tensor([float(last["iters"].max()), float(last["iters"].mean()), float(P_local), dev_s],
                            dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)  # device times: max over ranks
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        pr = [torch.empty_like(per_rank) for _ in range(world)]
        dist.all_gather(pr, per_rank)
        per_rank_table = torch.stack(pr).cpu().numpy()
        # the path's only collective: gather per-problem energies / iteration counts / flags
        if is_ensemble and job.capi_gather is not None:
            rec = job.capi_gather.gather(eng, P_total)              # P_total x (E, ||e||, n_iter, converged, status)
            ensemble_table = np.stack([rec[:, 0], rec[:, 2], rec[:, 3]], 1)
        else:
            local = ENS.pack_records(last["e"], eng.error_norm(), last["iters"], last["conv"], eng.status())
            if is_ensemble:
                rec = ENS.gather_records(local, P_total)
            else:  # replicas: one record per rank
                rec = ENS.gather_records(local, world)
            ensemble_table = np.stack([rec[:, 0], rec[:, 2], rec[:, 3]], 1)
    else:
        per_rank_table = per_rank.cpu().numpy()[None]
        ensemble_table = np.stack([last["e"], last["iters"].astype(float), last["conv"].astype(float)], 1)
    dev_s, e2e_s, t_total, fock_s = [float(x) for x in red.cpu()]
    iters_all, conv_all, launches_all = [float(x) for x in tot.cpu()]

    line = None
    if rank == 0:
        hbm_peak, peak_src = measured_peaks()
        ph = last["ph"]
        # dominant kernel: the fused one-launch Householder tridiagonalisation (trd_fused.cuh; phase id 11 brackets
        # its launches on the engine's stream)
        roofline = None
        n_eig = int(last["iters"].max())
        if ph["counts"].get("panel_kernel", 0) > 0:
            launches_k = ph["counts"]["panel_kernel"]
            avg_s = ph["panel_kernel"] / 1e3 / launches_k
            # algorithmic bytes: the product with the trailing matrix must read one triangle of it per column,
            # 4 (n-i-1)^2 bytes, from wherever the matrix lives.  For n <= ~1500 the kernel keeps
            # the matrix in registers / shared memory, so DRAM traffic (`traffic`, from ncu) is only ~8 n^2 bytes and the
            # kernel is bound by its one L2 exchange per column, not by bandwidth: see latency_model.
            total_bytes = sum(4.0 * (n - i - 1) ** 2 for i in range(n - 1)) * nblk * P_local
            per_launch = total_bytes * n_eig / launches_k  # n_eig tridiagonalisations in the solve loop
            ach = per_launch / avg_s / 1e9
            sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
            mats_in_flight = max(1, min(P_local * nblk, 148))
            if n <= 128:
                kname = ("trd_solo_reg_kernel<128> / <64> (single-CTA Householder tridiagonalisation, the matrix in "
                         "registers, one CTA per matrix of the batch; the trailing 64 x 64 block in a second launch)")
            elif P_local * nblk > 2:
                kname = ("trd_fused_kernel, symmetric-half variant (one cluster of 4 CTAs per matrix working in L2, two "
                         "CTAs per SM; rank-2 update fused with the trailing-matrix product on the lower triangle)")
            elif n > 1664 + 128:
                kname = ("trd_fused_blocked_kernel -> trd_fused_kernel (one-launch Householder tridiagonalisation on all "
                         "SMs: deferred updates while the trailing matrix exceeds L2, then in place in L2, then in shared "
                         "memory; one or two L2 exchanges per column)")
            else:
                kname = ("trd_fused_kernel (one-launch Householder tridiagonalisation: rank-2 update fused "
                         "with the trailing-matrix product, one L2 exchange per column)")
            roofline = {"kernel": kname, "bound": "hbm",
                        "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                        "traffic": ncu_traffic(n),
                        "peak_source": peak_src, "avg_launch_ms": avg_s * 1e3, "launches_per_solve": launches_k,
                        "share_of_scf_path": ph["panel_kernel"] / max(1e-9, last["scf_ms"]),
                        "latency_model": {
                            "what": "the matrix never leaves the SMs for n <= ~1500, so the HBM fraction measures the "
                                    "serial chain of n-1 column steps, not traffic; the floor of a column step is the "
                                    "cross-CTA all-gather of y through L2",
                            "cycles_per_column": avg_s * sm_mhz * 1e6 / max(1, n - 1) if P_local * nblk == 1 else None,
                            "exchange_floor_cycles": 2300,
                            "exchange_floor_source": "scripts/micro/exchange.cu on B200: ~4 dependent L2 round trips "
                                                     "(3600 cycles for 148 CTAs with fence + atomic + poll)",
                            "sm_mhz": sm_mhz, "matrices_in_flight": mats_in_flight}}
            if 192 < n <= 1664 + 128 and P_local * nblk <= 2:
                roofline["launch_note"] = ("one 'launch' = one tridiagonalisation = three kernels: the all-SM register / "
                                           "shared-memory kernel down to trailing order 128, then the single-CTA register "
                                           "kernels (128-row and 64-row layouts); cycles_per_column averages over all of them")
            if n > 1664 + 128 and P_local * nblk == 1:
                roofline["launch_note"] = ("one 'launch' = one tridiagonalisation = two kernels: the L2/HBM kernel down "
                                           "to trailing order 1664, then the shared-memory kernel (hand-over); "
                                           "traffic is the ncu figure of the first kernel only")
        flops_it = 7.0 * n ** 3 + 10.0 * n * n * na
        fp64_tf = flops_it * (iters_all / steps / world) * nblk / (dev_s / steps) / 1e12 if dev_s > 0 else 0.0
        # FP64 roofline of the whole iteration: LAPACK-equivalent flops W_it = 7 n^3 + 10 n^2 o per rank
        # against the DMMA peak measured on this GPU by a register-resident mma.sync.m8n8k4.f64 loop (outside the timed
        # region; MEASURED_PEAKS.json only has bf16 and HBM)
        dm, df, cp = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_double(0)
        fp64_roof = None
        if L.lib.gscf_b200_measure_peaks(ctypes.byref(dm), ctypes.byref(df), ctypes.byref(cp)) == L.OK and dm.value > 0:
            fp64_roof = {"bound": "tensor", "achieved": fp64_tf, "peak": dm.value, "unit": "TFLOP/s",
                         "frac": fp64_tf / dm.value, "dfma_peak": df.value,
                         "peak_source": "builder-measured live (gscf_b200_measure_peaks: register-resident "
                                        "mma.sync.m8n8k4.f64 loop on this GPU); MEASURED_PEAKS.json has no FP64 figure",
                         "flops_per_iteration": flops_it * nblk, "scope": "whole SCF iteration, per GPU"}
        line = {
            "metric": "scf_iterations_per_s", "value": iters_all / dev_s, "unit": "iterations/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * t_total / steps,
            "higher_is_better": True, "scaling": "strong" if is_ensemble else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["desc"], "problems_total": P_total if is_ensemble else world,
                       "problems_per_gpu": [int(x) for x in per_rank_table[:, 2]] if is_ensemble else 1,
                       "n_bf": n, "n_alpha": na, "n_beta": nb,
                       "n_aux": aux, "diis_depth": m, "iterations_per_solve": int(last["iters"].max()),
                       "parallelism": ("ensemble sharded by problem index (static contiguous blocks, no data-path "
                                       "collective)" if is_ensemble else
                                       "independent replica per GPU (single problem stays on one GPU)"),
                       "gather": job.gather_kind,
                       "timing": "value: CUDA-event device time of orthogonaliser set-up + solve loop, Fock builder "
                                 "excluded; e2e: wall clock with host S and guess coefficients in, C/eps/E out, Fock builder excluded",
                       "l2": "256 MiB buffer written between steps (L2 flush)",
                       "scf_ms_per_step": 1e3 * dev_s / steps, "fock_ms_per_step": 1e3 * fock_s / steps,
                       "phase_ms_last_step": {k: v for k, v in ph.items() if k != "counts"},
                       "iterations_per_rank_max_mean": [[float(r[0]), float(r[1])] for r in per_rank_table],
                       "scf_s_per_rank": [float(r[3]) for r in per_rank_table],
                       "fp64_useful_tflops": fp64_tf, "problems_converged_per_s": conv_all / dev_s,
                       "problems_converged_per_s_e2e": conv_all / e2e_s,
                       "converged": int(ensemble_table[:, 2].sum()), "records_gathered": int(ensemble_table.shape[0]),
                       "energies_last_step": [float(x) for x in ensemble_table[:4, 0]]},
            "clocks": clocks,
            "fp64_roofline": fp64_roof,
            "e2e": {"value": iters_all / e2e_s, "unit": "iterations/s",
                    "h2d_bytes_per_step": int((1 + nblk) * P_local * n * n * 8), "d2h_bytes_per_step": int(last["d2h"])},
            "gpu_launches": int(launches_all),
            "roofline": roofline,
        }
        if world == 1 and cpu_baseline:
            cores = os.cpu_count() or 1
            if is_ensemble:
                # an ensemble's CPU baseline is one member per core, not one member on all cores
                v, cores, nmem, its_cpu = oracle_time_ensemble(w, seeds[:2 * cores])
                line["cpu_baseline"] = {
                    "value": v, "unit": "iterations/s", "cores": cores, "kind": "port",
                    "sample": f"{nmem} members (seeds {seeds[0]}..{seeds[nmem - 1]}), full DIIS solves, one member per core "
                              f"with BLAS threads = 1 (oracle/: LAPACK dsygvd + dgemm); Fock builder excluded"}
            elif n >= 2048:
                # bounded sample: 2 DIIS iterations; the SCF path's work per iteration does not depend on n_aux (only the
                # excluded Fock builder does), so the sample problem is generated with n_aux = 4 to stay within seconds
                ws = dict(w)
                ws["aux"] = 4
                n_it, t_scf, t_fock, _ = oracle_time_solve(ws, seed=seeds[0], max_iter=2)
                line["cpu_baseline"] = {
                    "value": n_it / t_scf, "unit": "iterations/s", "cores": cores, "kind": "port",
                    "sample": f"{n_it} DIIS iterations of problem seed {seeds[0]} generated with n_aux = 4 (same SCF-path work "
                              f"per iteration) on the host (oracle/: LAPACK dsygvd + dgemm, all BLAS threads); Fock builder excluded"}
            else:
                n_it, t_scf, t_fock, e_ref = oracle_time_solve(w, seed=seeds[0])
                line["cpu_baseline"] = {
                    "value": n_it / t_scf, "unit": "iterations/s", "cores": cores, "kind": "port",
                    "sample": f"1 full DIIS solve ({n_it} iterations) of problem seed {seeds[0]} on the host "
                              f"(oracle/: LAPACK dsygvd + dgemm, all BLAS threads); Fock builder excluded",
                    "energy_abs_diff_vs_gpu": abs(e_ref - float(last["e"][0]))}
    eng.close()
    fock.close()
    del S_host, H_host, C_host, G_host
    torch.cuda.empty_cache()
    return line


def run_ours(args, w):
    job = Job()
    headline = measure(job, w, args.steps, args.warmup, cpu_baseline=not args.no_cpu_baseline, sample_clocks=True,
                       problems=args.problems)
    extra = {}
    if args.workload == "c2" and not args.no_workloads:
        # every other BASELINE configuration under the same clock (bounded: 2 warm-up + 3 timed solves each): the large
        # single problem on one GPU only (it never leaves one GPU), the ensembles SHARDED over all ranks at every N
        names = (["c3"] if job.world == 1 else []) + ["c4", "c5"]
        for name in names:
            try:
                rec = measure(job, WORKLOADS[name], 3, 2, cpu_baseline=not args.no_cpu_baseline)
            except Exception as exc:  # noqa: BLE001 - an extra workload must never take the headline down
                rec = {"error": repr(exc)} if job.rank == 0 else None
            if job.rank == 0:
                extra[name] = rec
    if job.rank == 0:
        if extra:
            keep = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "config", "e2e",
                    "gpu_launches", "roofline", "fp64_roofline", "cpu_baseline", "error")
            headline["workloads"] = {k: {kk: v[kk] for kk in keep if kk in v} for k, v in extra.items()}
            for k, v in headline["workloads"].items():
                if "config" in v and WORKLOADS[k]["P"] > 1:
                    v["metric"] = "problems_converged_per_s"
                    v["iterations_per_s"] = v["value"]
                    v["value"] = v["config"]["problems_converged_per_s"]
                    v["unit"] = "problems/s"
                    v["e2e"] = dict(v["e2e"], iterations_per_s=v["e2e"]["value"],
                                    value=v["config"]["problems_converged_per_s_e2e"], unit="problems/s")
        print(json.dumps(headline))
    job.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--problems", type=int, default=0, help="ensemble size override (c4/c5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--n-eigenpairs", type=int, default=0,
                    help="measurement aid: run the workload with ScfBaseKeys::n_eigenpairs = K < n_bf (default: all)")
    ap.add_argument("--no-workloads", action="store_true",
                    help="headline only: skip the extra BASELINE configurations (c3 / c4 / c5) of the default run")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = WORKLOADS[args.workload]
    if args.n_eigenpairs:
        w = dict(w, n_eigenpairs=args.n_eigenpairs,
                 desc=w["desc"] + f" [n_eigenpairs = {args.n_eigenpairs}: partial-spectrum eigensolves, not the BASELINE config]")
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
