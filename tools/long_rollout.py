#!/usr/bin/env python
"""Reset-free long rollouts with random actions: does anything get dropped (capacity overflow) or fail (non-finite state)?

    python tools/long_rollout.py --task sawyer_peg --envs 4096 --steps 10000

Prints one JSON line with the engine's work counters (`overflow_states == 0` over a 10k-step peg rollout)."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from earl_benchmark_b200.envs import kitchen, sawyer_door, sawyer_peg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="sawyer_peg")
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=10000)
    a = ap.parse_args()
    dev = "cuda:0"
    if a.task == "kitchen":
        env = kitchen.Kitchen(num_envs=a.envs, device=dev, seed=0)
        env.seed(0)
        nact = 9
    else:
        cls = sawyer_door.SawyerDoorV2 if a.task == "sawyer_door" else sawyer_peg.SawyerPegV2
        env = cls(reward_type="sparse", num_envs=a.envs, device=dev, seed=0)
        nact = 4
    env.reset()
    gen = torch.Generator(device=dev)
    gen.manual_seed(7)
    ring = torch.rand((64, a.envs, nact), generator=gen, device=dev, dtype=torch.float32) * 2 - 1
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    finite = True
    for t in range(a.steps):
        ob, r, d, info = env.step(ring[(t * 7) % 64])
        if t % 1000 == 999:
            finite = finite and bool(torch.isfinite(ob).all())
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    w = env.work_counters()
    print(json.dumps(dict(task=a.task, envs=a.envs, steps=a.steps, seconds=el, env_steps_per_s=a.envs * a.steps / el,
                          observations_finite=finite, work={k: int(v) for k, v in w.items()})), flush=True)


if __name__ == "__main__":
    main()
