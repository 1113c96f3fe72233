#!/usr/bin/env python
"""Free-space weld-tracking metric over the shipped Sawyer demonstrations (DIAGNOSTIC TOOL, needs /root/reference).

Compiles the Sawyer scenes from the reference MJCF with optional parameter overrides, replays the first K steps of every
door / peg episode (before any contact) through the fp64 checker and reports the rms / max hand error per axis against
the recorded hand trajectory, plus the rest-pose errors.  Used to look for structural differences, not to fit constants.
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tools"))
from earl_benchmark_b200 import demos  # noqa: E402
from earl_benchmark_b200.envs import sawyer_door, sawyer_peg  # noqa: E402
from oracle.engine import SawyerDoorOracle, SawyerPegOracle  # noqa: E402


def free_steps(task, d, s, en, kmax):
    """Number of leading steps of an episode before any contact: door = until the handle first moves (minus two steps of
    margin); peg = a fixed 8 steps (the peg itself falls 5 mm onto the table from step 0, so it cannot serve as a marker)."""
    obs, nobs = d["observations"], d["next_observations"]
    if task == "sawyer_peg":
        return min(8, kmax, en - s)
    dev = np.abs(nobs[s:en, 4:7] - obs[s, 4:7]).max(1)
    return max(0, min(int(np.argmax(dev > 1e-4)) - 2, kmax))


def metric(model_door, model_peg, kmax=16, verbose=False):
    out = {}
    for task, model, cls, ref0 in (("sawyer_door", model_door, SawyerDoorOracle, sawyer_door.initial_states[0][:3]),
                                   ("sawyer_peg", model_peg, SawyerPegOracle, sawyer_peg.initial_states[0][:3])):
        if model is None:
            continue
        o = cls(model)
        out[task + "_rest_mm"] = np.round((o.reset()[:3] - ref0) * 1000, 2)
        errs = []
        for which in ("forward", "reverse"):
            d = demos.load(task, which)
            obs, nobs, act = d["observations"], d["next_observations"], d["actions"]
            for s, en in demos.episodes(d):
                k = free_steps(task, d, s, en, kmax)
                o.goal = obs[s][7:14].astype(np.float64)
                if task == "sawyer_door":
                    ob0 = o.reset(door_angle=float(demos.door_angle_from_obs(obs[s])))
                else:
                    ob0 = o.reset(peg_pos=demos.peg_position_from_obs(obs[s]).astype(np.float64))
                off = ob0[:3] - obs[s][:3]      # rest-pose offset, removed so that only the RESPONSE is compared
                for t in range(s, s + k):
                    ob, _ = o.step(act[t])
                    errs.append(ob[:4] - np.concatenate([off, [0]]) - nobs[t][:4])
        e = np.array(errs)
        out[task + "_rms_mm"] = np.round(np.sqrt((e[:, :3] ** 2).mean(0)) * 1000, 2)
        out[task + "_max_mm"] = np.round(np.abs(e[:, :3]).max(0) * 1000, 2)
        out[task + "_grip_max"] = float(np.abs(e[:, 3]).max())
        out[task + "_n"] = len(e)
    return out


if __name__ == "__main__":
    import compile_models as CM
    print(metric(CM.sawyer_door(), CM.sawyer_peg()))
