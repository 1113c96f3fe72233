#!/usr/bin/env python
"""Episode-level replay of the shipped Sawyer demonstrations through the fp64 checker (TEST / DIAGNOSTIC TOOL).

For every episode of demonstrations/<task>/<forward|reverse>: reset, reconstruct the start state from obs[0], replay the
recorded actions open loop and report (i) whether the episode reaches success, (ii) the step at which it does against
the recorded success step, (iii) hand / object tracking error, plus the per-step sparse-reward agreement NEXT TO the
all-zeros predictor.  `evaluate(model_door, model_peg)` is what tests/test_engine_oracle.py asserts on.

    python tools/demo_eval.py [--tran-scale S]
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from earl_benchmark_b200 import demos  # noqa: E402
from earl_benchmark_b200.envs import sawyer_door, sawyer_peg  # noqa: E402
from earl_benchmark_b200.mjcf.compile import Model  # noqa: E402
from oracle.engine import SawyerDoorOracle, SawyerPegOracle  # noqa: E402


def replay(oracle, task, which, verbose=False, extra=3):
    d = demos.load(task, which)
    obs, nobs, act = d["observations"], d["next_observations"], d["actions"]
    rew = d["rewards"].ravel()
    out = []
    for s, en in demos.episodes(d):
        oracle.goal = obs[s][7:14].astype(np.float64)
        if task == "sawyer_door":
            ob = oracle.reset(door_angle=float(demos.door_angle_from_obs(obs[s])))
        else:
            ob = oracle.reset(peg_pos=demos.peg_position_from_obs(obs[s]).astype(np.float64))
        start_err = np.abs(ob[:7] - obs[s][:7])
        r, hand, obj = [], [], []
        for t in range(s, en):
            ob, rr = oracle.step(act[t])
            r.append(rr)
            hand.append(np.abs(ob[:3] - nobs[t][:3]).max())
            obj.append(np.abs(ob[4:7] - nobs[t][4:7]).max())
        r = np.array(r)
        first = np.nonzero(r)[0]
        # the recording ends ON its success step, so "success within +3 steps of the recording" can only be judged by
        # stepping on: the last recorded action is held for up to `extra` more steps (not counted in the per-step agreement)
        late = -1
        if len(first) == 0:
            for k in range(extra):
                _, rr = oracle.step(act[en - 1])
                if rr:
                    late = (en - s) + k
                    break
        out.append(dict(task=task, which=which, n=en - s, reward=r, demo_reward=rew[s:en], hand=np.array(hand), obj=np.array(obj),
                        success=len(first) > 0, step=int(first[0]) if len(first) else late, demo_step=int(np.nonzero(rew[s:en])[0][0]),
                        success_late=late >= 0, start_err=start_err))
        if verbose:
            e = out[-1]
            print(f"  {task} {which} ep{len(out) - 1}: n={e['n']} success={e['success']} late={e['success_late']} step={e['step']} demo_step={e['demo_step']} "
                  f"hand_max={e['hand'].max():.4f} obj_max={e['obj'].max():.4f} start_err={start_err.max():.4f}")
    return out


def summarise(eps):
    total = sum(e["n"] for e in eps)
    mism = sum(int((e["reward"] != e["demo_reward"]).sum()) for e in eps)
    zeros = sum(int((e["demo_reward"] != 0).sum()) for e in eps)
    return dict(episodes=len(eps), success=sum(e["success"] for e in eps),
                within3=sum((e["success"] or e["success_late"]) and abs(e["step"] - e["demo_step"]) <= 3 for e in eps),
                success_incl_3_more_steps=sum(e["success"] or e["success_late"] for e in eps),
                agreement=1 - mism / total, all_zeros=1 - zeros / total,
                hand_max=max(e["hand"].max() for e in eps), obj_max=max(e["obj"].max() for e in eps))


def evaluate(model_door=None, model_peg=None, verbose=False, tasks=("sawyer_door", "sawyer_peg")):
    res = {}
    if "sawyer_door" in tasks:
        o = SawyerDoorOracle(model_door or Model.load(sawyer_door.MODEL_PATH))
        rest = o.reset()[:3] - sawyer_door.initial_states[0][:3]
        res["door_rest_err"] = rest
        for w in ("forward", "reverse"):
            res[f"door_{w}"] = summarise(replay(o, "sawyer_door", w, verbose))
    if "sawyer_peg" in tasks:
        o = SawyerPegOracle(model_peg or Model.load(sawyer_peg.MODEL_PATH))
        rest = o.reset()[:3] - sawyer_peg.initial_states[0][:3]
        res["peg_rest_err"] = rest
        for w in ("forward", "reverse"):
            res[f"peg_{w}"] = summarise(replay(o, "sawyer_peg", w, verbose))
    return res


if __name__ == "__main__":
    scale = None
    if "--tran-scale" in sys.argv:
        scale = float(sys.argv[sys.argv.index("--tran-scale") + 1])
    md, mp = Model.load(sawyer_door.MODEL_PATH), Model.load(sawyer_peg.MODEL_PATH)
    if scale is not None:
        from earl_benchmark_b200.mjcf.compile import WELD_TRAN_SCALE
        md.weld_invweight[:, 0] *= scale / WELD_TRAN_SCALE
        mp.weld_invweight[:, 0] *= scale / WELD_TRAN_SCALE
    for k, v in evaluate(md, mp, verbose="-v" in sys.argv).items():
        print(k, v)
