#!/usr/bin/env python
"""Throughput of the batched kitchen task on one GPU (CUDA-event timing): reset, then random-action env steps."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from earl_benchmark_b200.envs import kitchen  # noqa: E402


def run(n, steps, warmup):
    env = kitchen.Kitchen(num_envs=n, device="cuda:0", seed=0)
    env.seed(0)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.rand((warmup + steps, n, 9), generator=g, device="cuda") * 2 - 1
    for t in range(warmup):
        env.step(a[t])
    torch.cuda.synchronize()
    w0 = env.work_counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(steps):
        env.step(a[warmup + t])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    w1 = env.work_counters()
    d = {k: w1[k] - w0[k] for k in w1}
    sub = max(1, d["substeps"])
    return dict(num_envs=n, steps=steps, ms_per_step=ms / steps, env_steps_per_s=n * steps / (ms * 1e-3),
                newton_per_substep=d["newton_iterations"] / sub, rows_per_substep=d["constraint_rows"] / sub,
                contacts_per_substep=d["contacts"] / sub, bad_states=d["bad_states"], overflow_states=d["overflow_states"])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, nargs="+", default=[740, 2960, 11840])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()
    for n in a.envs:
        print(json.dumps(run(n, a.steps, a.warmup)), flush=True)
