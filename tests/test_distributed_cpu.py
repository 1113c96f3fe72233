"""N>1 host logic on CPU: world_size-2 gloo processes (no GPU): shard ranges, goal-stream column slicing
and the one collective of the system (the eval-statistics all-reduce)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from earl_benchmark_b200 import rng, shard_range
from earl_benchmark_b200.distributed import all_reduce_eval_stats, max_over_ranks


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, rows, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    # this rank's slice of the global goal stream, exactly as TabletopManipulation._ensure builds it
    stream = rng.PyRandom(3).tabletop_goal_rows(rows * n_total).reshape(rows, n_total)[:, lo:hi]
    np.save(os.path.join(out_dir, f"stream_{rank}.npy"), stream)
    # per-rank eval statistics -> job-wide
    ret = float(np.arange(lo, hi).sum())
    stats = torch.tensor([ret, float(hi - lo) * 0.5, float(hi - lo) * 0.75, float(hi - lo)], dtype=torch.float64)
    res = all_reduce_eval_stats(stats)
    slow = max_over_ranks(10.0 + rank)
    # ranks sharing one host keep the staged copy pipeline of the host step (earl_set_host_zerocopy, envs/_hostio.py)
    from earl_benchmark_b200.envs import _hostio
    os.environ.pop("EARL_TT_HOST_ZEROCOPY", None)
    policy = _hostio.host_zerocopy_default()
    os.environ.pop("WORLD_SIZE")                       # no launcher variables: the initialised process group answers
    os.environ.pop("LOCAL_WORLD_SIZE", None)
    ranks = _hostio.ranks_on_this_host()
    if rank == 0:
        np.save(os.path.join(out_dir, "res.npy"), np.array([res["mean_return"], res["success_rate"], res["success_any_rate"],
                                                             res["num_envs"], slow, policy, ranks]))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    n_total, rows, world = 37, 5, 2
    mp.spawn(_worker, args=(world, _free_port(), n_total, rows, str(tmp_path)), nprocs=world, join=True)
    full = rng.PyRandom(3).tabletop_goal_rows(rows * n_total).reshape(rows, n_total)
    got = np.concatenate([np.load(tmp_path / f"stream_{r}.npy") for r in range(world)], axis=1)
    assert np.array_equal(got, full)  # sharded job draws the same goals as the single-process job
    res = np.load(tmp_path / "res.npy")
    assert res[3] == n_total and abs(res[0] - np.arange(n_total).sum() / n_total) < 1e-12
    assert abs(res[1] - 0.5) < 1e-12 and abs(res[2] - 0.75) < 1e-12 and res[4] == 11.0
    assert res[5] == 0 and res[6] == world             # two ranks on this host: staged pipeline for the host step


def test_single_process_degenerate_path():
    stats = torch.tensor([6.0, 1.0, 2.0, 4.0], dtype=torch.float64)
    res = all_reduce_eval_stats(stats)
    assert res == {"mean_return": 1.5, "success_rate": 0.25, "success_any_rate": 0.5, "num_envs": 4}
    assert max_over_ranks(3.5) == 3.5
