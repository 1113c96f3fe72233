#!/usr/bin/env python
"""Small kitchen run for compute-sanitizer (memcheck): reset (400 settle substeps) + env steps that drive the arm into the
cabinets, so the capsule / mesh / box narrow phase and the pyramidal rows run under the checker.
    compute-sanitizer --tool memcheck python tools/sanitize_kitchen.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from earl_benchmark_b200.envs import kitchen  # noqa: E402

n = 7
env = kitchen.Kitchen(num_envs=n, device="cuda:0", seed=1)
env.seed(1)
env.reset(config_index=np.arange(n) % 6)
a = torch.zeros((n, 9), device="cuda")
a[:, 0], a[:, 1], a[:, 2] = 0.9, 0.8, 0.9          # towards the upper cabinets
tot = 0
for t in range(int(sys.argv[1]) if len(sys.argv) > 1 else 25):
    ob, r, d, info = env.step(a)
w = env.work_counters()
print("kitchen sanitize run:", w, "finite", bool(torch.isfinite(ob).all()))
assert w["bad_states"] == 0
