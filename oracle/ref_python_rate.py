#!/usr/bin/env python
"""Throughput of the UNMODIFIED reference's own Python step loop on this container's host cores (CPU baseline (i) of
SURVEY.md 8(d)).  TEST / MEASUREMENT INFRASTRUCTURE ONLY; needs /root/reference, so it runs in the build container, not on
the GPU box -- the result is committed as profiles/r02/ref_python_loop_rate.json and quoted by bench.py's cpu_baseline note.

    python oracle/ref_python_rate.py [--steps 20000] [--procs 1 8]

The reference (`earl_benchmark.EARLEnvs('tabletop_manipulation', reward_type='sparse', train_horizon=200000)` ->
PersistentStateWrapper -> TabletopManipulation.step, earl_benchmark/envs/tabletop_manipulation.py:123-139) is imported from
/root/reference behind the no-op MuJoCo stand-in of oracle/fakes/ (the task is kinematic; the real mujoco_py loop adds two
discarded mj_forward calls per step, so this number FLATTERS the reference).  One environment per process, as the reference
has no vector env; every process steps `--steps` random actions after a reset.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("EARL_REFERENCE", "/root/reference")


def worker(seed, steps, q):
    sys.path.insert(0, os.path.join(HERE, "fakes"))
    sys.path.insert(0, REF)
    import random

    import numpy as np
    random.seed(seed)
    np.random.seed(seed)
    import earl_benchmark
    train, _ = earl_benchmark.EARLEnvs("tabletop_manipulation", reward_type="sparse", train_horizon=200000).get_envs()
    train.reset()
    acts = np.random.RandomState(seed).uniform(-1, 1, (steps, 3)).astype(np.float32)
    for a in acts[:200]:
        train.step(a)
    t0 = time.perf_counter()
    for a in acts:
        train.step(a)
    q.put(time.perf_counter() - t0)


def rate(procs, steps):
    q = mp.Queue()
    ps = [mp.Process(target=worker, args=(k, steps, q)) for k in range(procs)]
    t0 = time.perf_counter()
    for p in ps:
        p.start()
    el = [q.get() for _ in ps]
    for p in ps:
        p.join()
    wall = time.perf_counter() - t0
    return dict(procs=procs, steps_per_proc=steps, env_steps_per_s=procs * steps / max(el), per_proc_env_steps_per_s=steps / (sum(el) / len(el)),
                loop_seconds_max=max(el), wall_seconds_incl_import=wall)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--procs", type=int, nargs="+", default=[1, 8])
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "profiles", "r02", "ref_python_loop_rate.json"))
    a = ap.parse_args()
    if not os.path.isdir(REF):
        sys.exit(f"{REF} not found: this measurement only runs where the reference is present")
    import platform
    res = dict(what="reference's own PersistentStateWrapper(TabletopManipulation).step loop, sparse reward, one env per process, "
                    "behind the no-op MuJoCo stand-in (oracle/fakes)", host_cpus=os.cpu_count(), machine=platform.processor() or platform.machine(),
               python=platform.python_version(), runs=[rate(p, a.steps) for p in a.procs])
    print(json.dumps(res, indent=1))
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)
