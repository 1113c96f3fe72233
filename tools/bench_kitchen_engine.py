#!/usr/bin/env python
"""Throughput of the kitchen capacity set of the device engine: N environments x 40 substeps (= one KitchenV0.step worth
of physics) from contact-rich states of a scripted checker rollout, CUDA-event timing."""
import argparse
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
from earl_benchmark_b200.kitchen_engine import KitchenEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, nargs="+", default=[740, 2960, 11840])
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    import test_kitchen_engine_gpu as T
    _, states = T._states(50)
    eng = KitchenEngine("cuda:0")
    dev = eng.device
    for n in a.envs:
        pick = [states[i % len(states)] for i in range(n)]
        f32 = lambda k: torch.tensor(np.stack([s[k] for s in pick]), dtype=torch.float32, device=dev)  # noqa: E731
        q, v, w, c = f32(0), f32(1), f32(2), f32(4)
        mp = torch.tensor(np.stack([s[3] for s in pick]), dtype=torch.float64, device=dev)
        eng.substeps(q, v, w, mp, c, nsub=40)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        iters = 0
        for _ in range(a.steps):
            info = eng.substeps(q, v, w, mp, c, nsub=40)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        inf = info.cpu().numpy()
        print(json.dumps(dict(num_envs=n, ms_per_env_step=ms, env_steps_per_s=n / (ms * 1e-3), substeps_per_s=40 * n / (ms * 1e-3),
                              newton_per_substep=float(inf[:, 2].mean() / 40), rows=float(inf[:, 0].mean()), contacts=float(inf[:, 1].mean()),
                              flagged=int((inf[:, 3] != 0).sum()))), flush=True)


if __name__ == "__main__":
    main()
