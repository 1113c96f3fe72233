#!/usr/bin/env python
"""Throughput of the batched Sawyer door / peg step on one GPU (CUDA-event timing).  bench.py carries the judged line;
this is the sweep tool behind profiles/*/README.md."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from earl_benchmark_b200.envs import sawyer_door, sawyer_peg  # noqa: E402


def run(n, steps, warmup, ring=16, task="sawyer_door"):
    cls = sawyer_door.SawyerDoorV2 if task == "sawyer_door" else sawyer_peg.SawyerPegV2
    env = cls(reward_type="sparse", num_envs=n, device="cuda:0")
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(1234)
    actions = torch.rand((ring, n, 4), generator=g, device="cuda") * 2 - 1
    for t in range(warmup):
        env.step(actions[t % ring])
    torch.cuda.synchronize()
    w0 = env.work_counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(steps):
        env.step(actions[t % ring])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    w1 = env.work_counters()
    d = {k: w1[k] - w0[k] for k in w1}
    return dict(num_envs=n, steps=steps, ms_per_step=ms / steps, env_steps_per_s=n * steps / (ms * 1e-3),
                newton_per_substep=d["newton_iterations"] / max(1, d["substeps"]),
                rows_per_substep=d["constraint_rows"] / max(1, d["substeps"]),
                contacts_per_substep=d["contacts"] / max(1, d["substeps"]), bad_states=d["bad_states"],
                overflow_states=d["overflow_states"])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, nargs="+", default=[4096, 16384, 65536])
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--task", default="sawyer_door", choices=["sawyer_door", "sawyer_peg"])
    a = ap.parse_args()
    for n in a.envs:
        print(json.dumps(run(n, a.steps, a.warmup, task=a.task)), flush=True)
