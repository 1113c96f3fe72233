#!/usr/bin/env python
"""Small Sawyer door + peg run for compute-sanitizer (memcheck): reset, a few dozen steps incl. contact-rich ones."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from earl_benchmark_b200 import demos  # noqa: E402
from earl_benchmark_b200.envs import sawyer_door, sawyer_peg  # noqa: E402

fwd = demos.load("sawyer_door", "forward")
eps = demos.episodes(fwd)
n = 40
env = sawyer_door.SawyerDoorV2(num_envs=n, device="cuda:0")
env.reset(door_angle=np.resize(demos.door_angle_from_obs(fwd["observations"][[a for a, _ in eps]]), n))
for t in range(60):  # demo actions: the gripper reaches and pushes the handle (MPR + box-box contacts)
    a = np.stack([fwd["actions"][eps[i % 5][0] + t] for i in range(n)])
    env.step(torch.from_numpy(a).cuda())
print("door", env.work_counters())
peg = sawyer_peg.SawyerPegV2(reward_type="sparse", num_envs=n, device="cuda:0")
peg.reset()
rs = np.random.RandomState(0)
for t in range(40):
    peg.step(torch.from_numpy(np.clip(rs.uniform(-1, 1, (n, 4)) + [0, 0, -0.6, 0], -1, 1).astype(np.float32)).cuda())
print("peg", peg.work_counters())
torch.cuda.synchronize()
