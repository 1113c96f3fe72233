#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

TEST INFRASTRUCTURE ONLY.  Run in the build container, where /root/reference exists:

    python oracle/gen_golden.py

The reference package is imported from /root/reference behind the fake `gym` backend in
oracle/fakes/ (no MuJoCo needed: the tabletop task is kinematic, see oracle/fakes/gym/__init__.py).
Everything written here is an OUTPUT of the reference's own code
(`earl_benchmark/envs/tabletop_manipulation.py`, `earl_benchmark/wrappers/*.py`,
`earl_benchmark/__init__.py`) on seeded inputs, or a format conversion of the demonstration
pickles it ships.  /root/reference does not exist on the GPU box, so tests only read the
committed .npz files.

Interpreter note: this container has numpy 2.3 (NEP 50 scalar promotion); the reference pins
numpy 1.22.2.  The only observable difference on this path is the sparse-reward comparison
`float32_norm <= 0.2` (compared in fp32 here, in fp64 under 1.22) which differs only when the
fp32 norm equals float32(0.2) exactly, and the dense reward, evaluated in fp32 here and fp64
under 1.22 (agreement ~1e-7).  Both sit inside the tolerances the parity tests state.
"""
import os
import pickle
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("EARL_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "fakes"))
sys.path.insert(0, REF)

import earl_benchmark  # noqa: E402  (the reference)
from earl_benchmark.envs import tabletop_manipulation as ref_tt  # noqa: E402
from earl_benchmark.wrappers import lifelong_wrapper, persistent_state_wrapper  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")
DEMO_OUT = os.path.join(REPO, "earl_benchmark_b200", "demonstrations")


def scripted_actions(n, seed):
    """Random actions mixed with goal-seeking segments so attach / drag / clip / success all occur."""
    rs = np.random.RandomState(seed)
    a = rs.uniform(-1, 1, (n, 3)).astype(np.float32)
    # every other 256-step block: grip pinned on or off for long stretches
    for b in range(0, n, 256):
        if (b // 256) % 2 == 1:
            a[b:b + 256, 2] = np.where(rs.uniform(size=min(256, n - b)) < 0.9, 1.0, -1.0)
    a[::97] = np.float32(1.7) * a[::97]  # out-of-range actions exercise the input clip
    return a


def rollout(env, actions, policy_gain=None, reset_on_done=True, record_qpos=True):
    """Step `env` (a reference wrapper stack) and record everything observable."""
    n = len(actions)
    obs = np.zeros((n + 1, 12), np.float32)
    rew = np.zeros(n, np.float64)
    done = np.zeros(n, np.uint8)
    succ = np.zeros(n, np.uint8)
    qpos = np.zeros((n + 1, 4), np.float64)
    resets = np.zeros(n, np.uint8)  # 1 where the harness called reset() AFTER this step
    total = np.zeros(n, np.int64)
    interv = np.zeros(n, np.int64)
    used = np.zeros((n, 3), np.float32)
    o = env.reset()
    obs[0] = o
    qpos[0] = env.sim.data.qpos[:4]
    for t in range(n):
        a = actions[t].copy()
        if policy_gain is not None and (t // 512) % 2 == 0:
            # goal seeking: go to the mug, grip, drag it to the goal, release, return home
            fist, mug, goal = o[0:2], o[2:4], o[8:10]
            at_goal = np.linalg.norm(mug - goal) < 0.05
            if at_goal:
                tgt, grip = np.zeros(2, np.float32), -1.0
            elif o[4] == 0:
                tgt, grip = fist + (goal - mug), 1.0
            else:
                tgt, grip = mug, (1.0 if np.linalg.norm(fist - mug) < 0.3 else -1.0)
            a[:2] = np.clip(policy_gain * (tgt - fist) / 0.2 + 0.05 * a[:2], -1.5, 1.5)
            a[2] = grip
            a = a.astype(np.float32)
        used[t] = a
        o, r, d, _ = env.step(a)
        obs[t + 1] = o
        rew[t] = r
        done[t] = d
        succ[t] = bool(env.is_successful(o))
        total[t] = env.total_steps
        interv[t] = env.num_interventions
        qpos[t + 1] = env.sim.data.qpos[:4]
        if d and reset_on_done:
            o = env.reset()
            resets[t] = 1
            obs[t + 1] = o  # what the user sees next
            qpos[t + 1] = env.sim.data.qpos[:4]
    return dict(actions=used, obs=obs, reward=rew, done=done, success=succ, qpos=qpos,
                reset_after=resets, total_steps=total, num_interventions=interv)


def make_envs(seed, **kw):
    random.seed(seed)
    np.random.seed(seed)
    return earl_benchmark.EARLEnvs("tabletop_manipulation", **kw)


def gen_rollouts():
    out = {}
    # config 1 in miniature: sparse, train horizon 1000 so the reset-free horizon fires 4x
    loader = make_envs(0, reward_type="sparse", train_horizon=1000, eval_horizon=200)
    train, evl = loader.get_envs()
    n = 4096
    r = rollout(train, scripted_actions(n, 0), policy_gain=0.8)
    out.update({f"sparse_train_{k}": v for k, v in r.items()})
    r = rollout(evl, scripted_actions(1000, 1), policy_gain=0.8)
    out.update({f"sparse_eval_{k}": v for k, v in r.items()})
    # dense reward
    loader = make_envs(1, reward_type="dense", train_horizon=700, eval_horizon=200)
    train, _ = loader.get_envs()
    r = rollout(train, scripted_actions(2048, 2), policy_gain=0.8)
    out.update({f"dense_train_{k}": v for k, v in r.items()})
    # reset_train_env_at_goal
    loader = make_envs(2, reward_type="sparse", reset_train_env_at_goal=True, train_horizon=300)
    train, _ = loader.get_envs()
    r = rollout(train, scripted_actions(1500, 3), policy_gain=0.8)
    out.update({f"resetgoal_train_{k}": v for k, v in r.items()})
    # wide_init_distr (np.random-driven rejection-sampled resets; success on the mug only)
    loader = make_envs(3, reward_type="sparse", wide_init_distr=True, train_horizon=250)
    train, _ = loader.get_envs()
    r = rollout(train, scripted_actions(1500, 4), policy_gain=0.8)
    out.update({f"wide_train_{k}": v for k, v in r.items()})
    # no reset after done: counters keep growing, done stays True
    loader = make_envs(4, reward_type="sparse", train_horizon=100)
    train, _ = loader.get_envs()
    r = rollout(train, scripted_actions(300, 5), reset_on_done=False)
    out.update({f"noreset_train_{k}": v for k, v in r.items()})
    np.savez_compressed(os.path.join(GOLD, "tabletop_ref_rollouts.npz"), **out)
    print("rollouts:", {k: v.shape for k, v in out.items() if k.endswith("_obs")},
          "successes", {k: int(v.sum()) for k, v in out.items() if k.endswith("_success")},
          "attached rows", int((out["sparse_train_obs"][:, 4] == 0).sum()))


def gen_lifelong():
    random.seed(7)
    np.random.seed(7)
    loader = earl_benchmark.EARLEnvs("tabletop_manipulation", reward_type="sparse",
                                     setup_as_lifelong_learning=True, train_horizon=400,
                                     goal_change_frequency=37)
    env = loader.get_envs()
    assert isinstance(env, lifelong_wrapper.LifelongWrapper)
    n = 1500
    acts = scripted_actions(n, 6)
    obs = np.zeros((n + 1, 12), np.float32)
    rew = np.zeros(n)
    done = np.zeros(n, np.uint8)
    ll = np.zeros(n)
    used = np.zeros((n, 3), np.float32)
    resets = np.zeros(n, np.uint8)
    interv = np.zeros(n, np.int64)
    o = env.reset()
    obs[0] = o
    for t in range(n):
        a = acts[t].copy()
        if (t // 300) % 2 == 0:
            fist, mug, goal = o[0:2], o[2:4], o[8:10]
            if o[4] == 0:
                tgt, grip = fist + (goal - mug), 1.0
            else:
                tgt, grip = mug, (1.0 if np.linalg.norm(fist - mug) < 0.3 else -1.0)
            if np.linalg.norm(mug - goal) < 0.05:
                tgt, grip = np.zeros(2, np.float32), -1.0
            a[:2] = np.clip(0.8 * (tgt - fist) / 0.2, -1, 1)
            a[2] = grip
            a = a.astype(np.float32)
        used[t] = a
        o, r, d, _ = env.step(a)
        obs[t + 1], rew[t], done[t], ll[t] = o, r, d, env.lifelong_return
        interv[t] = env.num_interventions
        if d:
            o = env.reset()
            obs[t + 1] = o
            resets[t] = 1
    np.savez_compressed(os.path.join(GOLD, "tabletop_ref_lifelong.npz"), actions=used, obs=obs, reward=rew,
                        done=done, lifelong_return=ll, reset_after=resets, num_interventions=interv,
                        goal_change_frequency=37, train_horizon=400, seed=7)
    print("lifelong: return", ll[-1], "goal changes", int((np.abs(np.diff(obs[:, 8:10], axis=0)).sum(1) > 0).sum()))


def gen_goal_streams():
    """Goal-row index streams from the reference's own get_next_goal() (global `random`)."""
    out = {}
    goal_xy = ref_tt.goal_states[:, 2:4]
    for seed in (0, 1, 123, 2**31 + 5, 2**40 + 17):
        random.seed(seed)
        env = ref_tt.TabletopManipulation(reward_type="sparse")
        random.seed(seed)  # construction draws nothing, but make the stream origin explicit
        rows = []
        for _ in range(256):
            g = env.get_next_goal()
            rows.append(int(np.argmin(np.abs(goal_xy - g[2:4]).sum(1))))
        out[f"seed_{seed}"] = np.array(rows, np.int32)
    # N reference envs in one process sharing the global stream, reset in env order, 3 rounds
    random.seed(11)
    envs = [persistent_state_wrapper.PersistentStateWrapper(ref_tt.TabletopManipulation(reward_type="sparse"), 50)
            for _ in range(16)]
    random.seed(11)
    rounds = []
    for _ in range(3):
        rounds.append([int(np.argmin(np.abs(goal_xy - e.reset()[8:10]).sum(1))) for e in envs])
    out["shared_stream_seed11_16envs_3resets"] = np.array(rounds, np.int32)
    # wide-init states: global np.random uniform + rejection (reference reset(), :114-117)
    np.random.seed(5)
    random.seed(5)
    env = ref_tt.TabletopManipulation(reward_type="sparse", wide_init_distr=True)
    np.random.seed(5)
    random.seed(5)
    out["wide_init_seed5"] = np.stack([env.reset()[:4].astype(np.float64) for _ in range(64)])
    np.random.seed(5)
    wide64 = []
    for _ in range(64):
        env.reset()
        wide64.append(env.sim.data.qpos[:4].copy())
    out["wide_init_seed5_f64"] = np.stack(wide64)
    np.savez_compressed(os.path.join(GOLD, "tabletop_ref_streams.npz"), **out)
    print("goal stream seed 0:", out["seed_0"][:12])


def gen_demo_replay():
    """Per-transition replay of the shipped tabletop demos through the reference step()."""
    loader = make_envs(0, reward_type="sparse")
    fwd, rev = loader.get_demonstrations()
    out = {}
    for name, demo in (("forward", fwd), ("reverse", rev)):
        train, _ = make_envs(0, reward_type="sparse").get_envs()
        train.reset()
        env = train.env
        n = len(demo["actions"])
        ob = np.zeros((n, 12), np.float32)
        rw = np.zeros(n)
        for t in range(n):
            o = demo["observations"][t]
            env.attached_object = (0, 0) if o[4] == 0 else (-1, -1)
            env.set_state(o[:4].astype(np.float64))
            env.goal = o[6:].astype(np.float64)
            ob[t], rw[t], d, _ = train.step(demo["actions"][t])
        out[f"{name}_ref_next_obs"] = ob
        out[f"{name}_ref_reward"] = rw
        err = np.abs(ob - demo["next_observations"]).max()
        mism = int((rw != demo["rewards"][:, 0]).sum())
        print(f"demo {name}: reference-vs-stored max|obs err| {err:.3e}, reward mismatches {mism}")
    np.savez_compressed(os.path.join(GOLD, "tabletop_ref_demo_replay.npz"), **out)


def convert_demos():
    """Re-encode the shipped demonstration pickles as .npz package data (same arrays, same keys)."""
    for env in ("tabletop_manipulation", "sawyer_door", "sawyer_peg"):
        for direction in ("forward", "reverse"):
            src = os.path.join(REF, "earl_benchmark", "demonstrations", env, direction, "demo_data.pkl")
            with open(src, "rb") as f:
                d = pickle.load(f)
            dst_dir = os.path.join(DEMO_OUT, env)
            os.makedirs(dst_dir, exist_ok=True)
            np.savez_compressed(os.path.join(dst_dir, f"{direction}.npz"), **{k: np.asarray(v) for k, v in d.items()})


def gen_loader_constants():
    out = {}
    for env in ("tabletop_manipulation", "sawyer_door", "sawyer_peg"):
        mod = __import__(f"earl_benchmark.envs.{env}", fromlist=["x"]) if env == "tabletop_manipulation" else None
        if mod is not None:
            out[f"{env}_initial_states"] = mod.initial_states
            out[f"{env}_goal_states"] = mod.goal_states
    np.savez_compressed(os.path.join(GOLD, "loader_constants.npz"), **out)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    gen_rollouts()
    gen_lifelong()
    gen_goal_streams()
    gen_demo_replay()
    convert_demos()
    gen_loader_constants()
    print("done")
