#!/usr/bin/env python
"""Generate tests/golden/kitchen_ref_logic.npz by RUNNING THE UNMODIFIED REFERENCE kitchen task logic.

TEST INFRASTRUCTURE ONLY.  Run in the build container, where /root/reference exists:

    python oracle/gen_kitchen_golden.py

`earl_benchmark/envs/kitchen.py::Kitchen` (-> adept_envs KitchenV0 / RobotEnv / MujocoEnv / Robot_VelAct) is imported
from /root/reference behind the stand-ins in oracle/fakes/ (gym 0.23.1 seeding restated, termcolor, and a `mujoco_py`
whose MjSim.step() is a SCRIPTED pseudo-dynamics: MuJoCo is not in this image).  What is recorded is therefore what
the reference's own Python computes AROUND the physics, on states we feed it:
  * KitchenV0.step: clip / scale of the action, mocap update and clipping, Robot_VelAct control from the LAST NOISY
    observation, position clipping, the number of sim.step() calls and the ctrl vector they see   (glue)
  * Robot.get_obs / KitchenV0._get_obs: observation noise drawn from env.np_random                  (row a13)
  * Kitchen._get_reward_n_score, is_successful                                                       (row a14)
  * Kitchen.reset_model: np.random.randint over the six initial configurations, 10 settle steps     (row a11, kitchen)
The scripted states and site positions are inputs, not claims about MuJoCo.
Interpreter note: Python 3.12 / numpy 2.3 here vs 3.7 / 1.22.2 pinned by the reference; `collections.Mapping` is
aliased for the import.  All recorded arithmetic is float64 numpy ufuncs and `np.linalg.norm`.
"""
import collections
import collections.abc
import os
import sys

import numpy as np

collections.Mapping = collections.abc.Mapping
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("EARL_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "fakes"))
sys.path.insert(0, REF)

from earl_benchmark.envs import kitchen as ref_kitchen  # noqa: E402  (the reference)
import mujoco_py  # noqa: E402  (the stand-in)

OUT = os.path.join(REPO, "tests", "golden", "kitchen_ref_logic.npz")


def sites(env):
    return np.stack([env.sim.data.get_site_xpos(s) for s in mujoco_py.SITES])


def episode(env, env_seed, np_seed, actions):
    env.seed(env_seed)
    np.random.seed(np_seed)
    n0 = env.sim.nsteps
    ob0 = env.reset()
    rec = dict(reset_obs=ob0, reset_qpos=env.sim.data.qpos.copy(), reset_mocap=env.sim.data.mocap_pos[0].copy(),
               reset_sim_steps=env.sim.nsteps - n0, reset_ctrl=np.array(env.sim.ctrl_log[-1]),
               reset_success=bool(env.is_successful(ob0)), reset_next_np_random=np.random.randint(1 << 30))
    keys = ("qpos", "mocap", "sites", "obs", "reward", "success", "ctrl", "sim_steps", "mocap_before", "qpos_before")
    out = {k: [] for k in keys}
    for a in actions:
        n0 = env.sim.nsteps
        out["mocap_before"].append(env.sim.data.mocap_pos[0].copy())
        out["qpos_before"].append(env.sim.data.qpos.copy())
        ob, r, done, info = env.step(a)
        assert done is False
        out["qpos"].append(env.sim.data.qpos.copy())
        out["mocap"].append(env.sim.data.mocap_pos[0].copy())
        out["sites"].append(sites(env))
        out["obs"].append(ob)
        out["reward"].append(r)
        out["success"].append(bool(env.is_successful(ob)))
        out["ctrl"].append(np.array(env.sim.ctrl_log[-1]))
        out["sim_steps"].append(env.sim.nsteps - n0)
    rec.update({k: np.array(v) for k, v in out.items()})
    return rec


def main():
    env = ref_kitchen.Kitchen()
    rs = np.random.RandomState(2024)
    data = dict(goal_states=ref_kitchen.goal_states, all_pairs=ref_kitchen.initial_states["all_pairs"],
                init_qpos=env.init_qpos, pos_noise_amp=env.robot.robot_pos_noise_amp, vel_noise_amp=env.robot.robot_vel_noise_amp,
                pos_bound=env.robot.robot_pos_bound, vel_bound=env.robot.robot_vel_bound, midpoint_pos=env.midpoint_pos,
                mocap_clip=np.stack([env.mocap_pos_clip_lower, env.mocap_pos_clip_upper]), act_amp=env.act_amp,
                site_names=np.array(mujoco_py.SITES), frame_skip=env.frame_skip, noise_ratio=env.robot_noise_ratio)
    n_ep, T = 6, 60
    for e in range(n_ep):
        a = rs.uniform(-1.3, 1.3, (T, 9))          # some entries outside [-1, 1]: the input clip
        a[::7, :3] = np.sign(a[::7, :3])           # saturated mocap moves run into the mocap clip box
        rec = episode(env, env_seed=100 + e, np_seed=7 + e, actions=a)
        for k, v in rec.items():
            data[f"ep{e}_{k}"] = v
        data[f"ep{e}_actions"] = a
        data[f"ep{e}_seeds"] = np.array([100 + e, 7 + e])
    # a goal-reaching state: objects at the goal, so the +1-per-component branch and success both fire
    env.seed(5)
    np.random.seed(5)
    env.reset()
    q = env.sim.data.qpos
    q[9:] = ref_kitchen.goal_states[0][9:] + 1e-4
    ob = env._get_obs()
    data["near_goal_obs"] = ob
    data["near_goal_mocap"] = env.sim.data.mocap_pos[0].copy()
    data["near_goal_sites"] = sites(env)
    data["near_goal_reward"] = env.compute_reward(ob)
    data["near_goal_success"] = bool(env.is_successful(ob))
    data["n_episodes"], data["T"] = n_ep, T
    np.savez_compressed(OUT, **data)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
