#!/usr/bin/env python
"""Kitchen throughput under a CONTACT-RICH scripted policy (all arms reach for the slide-cabinet handle and keep pushing into
it, plus action noise), as opposed to bench.py's random actions that rarely touch anything: env-steps/s, rows and contacts
per substep, share of env steps re-stepped by the large capacity set."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from earl_benchmark_b200.envs import kitchen  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4736)
    ap.add_argument("--warmup", type=int, default=60)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    env = kitchen.Kitchen(num_envs=a.envs, device="cuda:0", seed=0)
    env.seed(0)
    env.reset()
    k = kitchen.REWARD_SITES.index("slide_site") if "slide_site" in kitchen.REWARD_SITES else 5
    gen = torch.Generator(device="cuda").manual_seed(3)

    def act():
        st = env.get_state()
        d = torch.from_numpy(st["site_xpos"][:, k] - st["mocap_pos"]).to("cuda", torch.float32)
        u = torch.zeros((a.envs, 9), device="cuda")
        u[:, :3] = torch.clamp(d * 10, -1, 1) * 0.5
        u += 0.1 * (torch.rand((a.envs, 9), generator=gen, device="cuda") * 2 - 1)
        return u

    for t in range(a.warmup):
        env.step(act())
    acts = [act() for _ in range(1)]
    torch.cuda.synchronize()
    w0 = env.work_counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(a.steps):
        env.step(acts[0])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    w1 = env.work_counters()
    d = {key: w1[key] - w0[key] for key in w1}
    sub = max(1, d["substeps"])
    print(json.dumps(dict(num_envs=a.envs, steps=a.steps, ms_per_step=ms / a.steps, env_steps_per_s=a.envs * a.steps / (ms * 1e-3),
                          rows_per_substep=d["constraint_rows"] / sub, contacts_per_substep=d["contacts"] / sub,
                          newton_per_substep=d["newton_iterations"] / sub, redone_share=d["redone_states"] / max(1, d["env_steps"]),
                          overflow_states=d["overflow_states"], bad_states=d["bad_states"])), flush=True)


if __name__ == "__main__":
    main()
