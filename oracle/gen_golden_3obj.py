#!/usr/bin/env python
"""Generate tests/golden/tabletop3_ref_rollouts.npz by RUNNING THE UNMODIFIED REFERENCE class
`earl_benchmark/envs/tabletop_manipulation_3obj.py` under its own PersistentStateWrapper.

TEST INFRASTRUCTURE ONLY.  Run in the build container, where /root/reference exists:

    python oracle/gen_golden_3obj.py

Same arrangement as oracle/gen_golden.py: the reference is imported from /root/reference behind the no-op
MuJoCo stand-in of oracle/fakes (the task is kinematic).  Interpreter note: numpy 2.3 here vs the reference's
numpy 1.22.2 pin -- the `<= 0.4` comparison and the dense reward are evaluated in fp32 here (fp64 there); the
fixture stores the fp32 norm-derived quantities, and the parity tests state the tolerances.
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("EARL_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "fakes"))
sys.path.insert(0, REF)

from earl_benchmark.envs import tabletop_manipulation_3obj as ref3  # noqa: E402  (the reference)
from earl_benchmark.wrappers import persistent_state_wrapper  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")
MARK = {(-1, -1): 0, (0, 0): 1, (0.5, 0.5): 2, (1, 1): 3}


def policy(o, a, t):
    """goal seeking in alternate 400-step blocks: fetch the first misplaced object, drag it to its goal"""
    if (t // 400) % 2 == 1:
        return a
    fist = o[0:2]
    for k in range(3):
        obj, goal = o[2 + 2 * k:4 + 2 * k], o[12 + 2 * k:14 + 2 * k]
        if np.linalg.norm(obj - goal) >= 0.05:
            break
    else:
        tgt, grip = o[10:12], -1.0
        a[:2] = np.clip(0.8 * (tgt - fist) / 0.2 + 0.05 * a[:2], -1.5, 1.5)
        a[2] = grip
        return a.astype(np.float32)
    held = o[8] == 0.5 * k
    if held:
        tgt, grip = fist + (goal - obj), 1.0
    else:
        tgt = obj
        grip = 1.0 if (np.linalg.norm(fist - obj) < 0.3 and o[8] == -1) else -1.0
    a[:2] = np.clip(0.8 * (tgt - fist) / 0.2 + 0.05 * a[:2], -1.5, 1.5)
    a[2] = grip
    return a.astype(np.float32)


def rollout(env, n, seed):
    rs = np.random.RandomState(seed)
    acts = rs.uniform(-1, 1, (n, 3)).astype(np.float32)
    acts[::97] *= np.float32(1.7)
    out = dict(actions=np.zeros((n, 3), np.float32), obs=np.zeros((n + 1, 20), np.float32), reward=np.zeros(n),
               done=np.zeros(n, np.uint8), success=np.zeros(n, np.uint8), qpos=np.zeros((n + 1, 8)),
               attached=np.zeros(n + 1, np.int32), reset_after=np.zeros(n, np.uint8),
               total_steps=np.zeros(n, np.int64), num_interventions=np.zeros(n, np.int64))
    o = env.reset()
    out["obs"][0], out["qpos"][0] = o, env.sim.data.qpos[:8]
    for t in range(n):
        a = policy(o, acts[t].copy(), t)
        out["actions"][t] = a
        o, r, d, _ = env.step(a)
        out["obs"][t + 1], out["reward"][t], out["done"][t] = o, r, d
        out["success"][t] = bool(env.is_successful(o))
        out["total_steps"][t], out["num_interventions"][t] = env.total_steps, env.num_interventions
        out["qpos"][t + 1], out["attached"][t + 1] = env.sim.data.qpos[:8], MARK[tuple(env.attached_object)]
        if d:
            o = env.reset()
            out["reset_after"][t] = 1
            out["obs"][t + 1], out["qpos"][t + 1], out["attached"][t + 1] = o, env.sim.data.qpos[:8], 0
    return out


def gen_attach_cases(n=4000):
    """one step from states with several objects near the fist: closest-object choice (3OBJ:101-108)"""
    rs = np.random.RandomState(7)
    env = ref3.TabletopManipulation(reward_type="sparse")
    env.reset()
    q0 = np.zeros((n, 8))
    acts = rs.uniform(-1, 1, (n, 3)).astype(np.float32)
    acts[: n // 2, 2] = np.abs(acts[: n // 2, 2])
    obs = np.zeros((n, 20), np.float32)
    q1 = np.zeros((n, 8))
    att = np.zeros(n, np.int32)
    for i in range(n):
        f = rs.uniform(-2.9, 2.9, 2)
        q = np.concatenate([f] + [f + rs.uniform(-0.45, 0.45, 2) for _ in range(3)])
        if i % 50 == 0:          # exact ties: two objects mirrored about the fist
            q[4:6] = f - (q[2:4] - f)
        full = np.zeros(9)
        full[:8], full[8] = q, -10
        env.set_state(full, env.sim.data.qvel.copy())
        env.attached_object = (-1, -1)
        q0[i] = env.sim.data.qpos[:8]
        o, _, _, _ = env.step(acts[i])
        obs[i], q1[i], att[i] = o, env.sim.data.qpos[:8], MARK[tuple(env.attached_object)]
    return dict(attach_q0=q0, attach_actions=acts, attach_obs=obs, attach_q1=q1, attach_att=att)


def main():
    out = {}
    random.seed(0)
    np.random.seed(0)
    env = persistent_state_wrapper.PersistentStateWrapper(ref3.TabletopManipulation(reward_type="sparse"), episode_horizon=900)
    out.update({f"sparse_{k}": v for k, v in rollout(env, 4000, 0).items()})
    np.random.seed(1)
    env = persistent_state_wrapper.PersistentStateWrapper(ref3.TabletopManipulation(reward_type="dense"), episode_horizon=700)
    out.update({f"dense_{k}": v for k, v in rollout(env, 2400, 1).items()})
    # reset_at_goal: every reset consumes np.random.randint(1) (nothing) + uniform(-0.3, 0.3, 8)
    np.random.seed(5)
    env = persistent_state_wrapper.PersistentStateWrapper(
        ref3.TabletopManipulation(reward_type="sparse", reset_at_goal=True), episode_horizon=50)
    out.update({f"rag_{k}": v for k, v in rollout(env, 400, 2).items()})
    out.update(gen_attach_cases())
    out["initial_states"], out["goal_states"] = ref3.initial_states, ref3.goal_states
    path = os.path.join(GOLD, "tabletop3_ref_rollouts.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;",
          "sparse successes", int(out["sparse_success"].sum()), "attach histogram", np.bincount(out["attach_att"], minlength=4),
          "rollout attach histogram", np.bincount(out["sparse_attached"], minlength=4))


if __name__ == "__main__":
    main()
