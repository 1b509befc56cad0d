"""world_size-2 gloo test of the N > 1 host logic (sharding + the single gather collective)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_problems, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from gscf_b200 import ensemble

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = ensemble.shard_range(n_problems, rank, world)
    idx = np.arange(lo, hi)
    # fake per-problem results that encode the problem index
    local = ensemble.pack_records(-10.0 - idx, 1e-7 * (idx + 1), 5 + idx % 3, (idx % 4 != 0), np.zeros_like(idx))
    rec = ensemble.gather_records(local, n_problems)
    if rank == 0:
        np.save(out, rec)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_problems", [7, 16])
def test_shard_and_gather_world2(tmp_path, n_problems):
    import torch.multiprocessing as mp

    out = str(tmp_path / "rec.npy")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_problems, out), nprocs=2, join=True)
    rec = np.load(out)
    idx = np.arange(n_problems)
    assert rec.shape == (n_problems, 5)
    np.testing.assert_array_equal(rec[:, 0], -10.0 - idx)
    np.testing.assert_array_equal(rec[:, 2], 5 + idx % 3)
    np.testing.assert_array_equal(rec[:, 3], (idx % 4 != 0).astype(float))


def test_shard_range_partitions():
    from gscf_b200.ensemble import shard_range, summarise

    for P in [1, 5, 8, 4096, 255]:
        for R in [1, 2, 4, 8]:
            blocks = [shard_range(P, r, R) for r in range(R)]
            assert blocks[0][0] == 0 and blocks[-1][1] == P
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(R - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    s = summarise(np.array([[-1.0, 1e-8, 5, 1, 0], [-2.0, 1e-3, 100, 0, 5]]))
    assert s == {"n_problems": 2, "n_converged": 1, "iterations_total": 105, "iterations_max": 100,
                 "iterations_mean": 52.5}


def _worker_solve(rank, world, port, n_problems, out, fail_rank):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from gscf_b200 import ensemble

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def fake(seeds):
        if rank == fail_rank:
            raise RuntimeError("local solve failed")
        idx = np.asarray(seeds)
        return ensemble.pack_records(-1.0 * idx, 1e-8 + 0 * idx, 7 + idx, np.ones_like(idx), np.zeros_like(idx))

    raised = False
    try:
        rec = ensemble.solve_ensemble(8, 1, n_problems=n_problems, local_solver=fake)
        if rank == 0:
            np.save(out, rec)
    except RuntimeError:
        raised = True
    assert raised == (rank == fail_rank)
    dist.barrier()  # nobody is stuck in the gather
    dist.destroy_process_group()


@pytest.mark.parametrize("n_problems,fail_rank", [(1, -1), (5, 1)])
def test_solve_ensemble_empty_shard_and_failing_rank(tmp_path, n_problems, fail_rank):
    """More ranks than problems (rank 1 owns nothing) and a rank whose local solve raises: every rank still reaches the
    single collective (the other ranks used to hang in all_gather)."""
    import torch.multiprocessing as mp

    out = str(tmp_path / "rec.npy")
    mp.spawn(_worker_solve, args=(2, _free_port(), n_problems, out, fail_rank), nprocs=2, join=True)
    rec = np.load(out)
    assert rec.shape == (n_problems, 5)
    if fail_rank < 0:
        np.testing.assert_array_equal(rec[:, 2], [7.0])
    else:
        np.testing.assert_array_equal(rec[:3, 2], [7.0, 8.0, 9.0])  # rank 0's shard
        assert np.all(rec[3:, 4] == -1.0)                              # rank 1 reported its failure
