"""Pin the CPU oracle (oracle/tabletop_oracle.c) against outputs of the UNMODIFIED reference.

Fixtures: tests/golden/tabletop_ref_*.npz, produced by oracle/gen_golden.py from
/root/reference/earl_benchmark/{envs/tabletop_manipulation.py,wrappers/*.py,__init__.py}, plus the
shipped demonstration transitions.
"""
import os

import numpy as np
import pytest

from oracle import loader
from oracle.loader import GOAL_STATES, TabletopOracle

BAND = 1e-5  # north_star: sparse reward bit-exact except within 1e-5 of the threshold


def goal_row(obs):
    d = np.abs(GOAL_STATES[None, :, 2:4] - obs[:, None, 8:10]).sum(-1)
    return d.argmin(1)


def success_norm(obs, wide):
    if wide:
        return np.linalg.norm(obs[:, 2:4].astype(np.float64) - obs[:, 8:10], axis=1)
    return np.linalg.norm(obs[:, :4].astype(np.float64) - obs[:, 6:10], axis=1)


def replay(gold, prefix, horizon, dense=False, wide=False, custom_init=False):
    g = {k[len(prefix) + 1:]: gold[k] for k in gold.files if k.startswith(prefix + "_")}
    n = len(g["actions"])
    orc = TabletopOracle(1, horizon, dense=dense, wide=wide)
    init = g["qpos"][0:1] if custom_init else None
    o = orc.reset(goal_row(g["obs"][0:1]), init_qpos=init)
    assert np.array_equal(o[0], g["obs"][0])
    for t in range(n):
        ob, rw, dn, sc = orc.step(g["actions"][t][None])
        # what the user sees after the optional reset is checked below; the step output itself:
        if not g["reset_after"][t]:
            assert np.array_equal(ob[0], g["obs"][t + 1]), (prefix, t)
            assert np.array_equal(orc.qpos[0], g["qpos"][t + 1]), (prefix, t)  # fp64 state, bit-exact
        nrm = success_norm(ob, wide)[0]
        if abs(nrm - 0.2) > BAND:
            assert bool(sc[0]) == bool(g["success"][t]), (prefix, t)
            if not dense:
                assert rw[0] == g["reward"][t], (prefix, t)
        if dense:
            assert abs(rw[0] - g["reward"][t]) <= 2e-6 * max(1.0, abs(rw[0])), (prefix, t)
        assert dn[0] == g["done"][t], (prefix, t)
        assert orc.total_steps[0] == g["total_steps"][t]
        if g["reset_after"][t]:
            nxt = g["obs"][t + 1:t + 2]
            init = g["qpos"][t + 1:t + 2] if custom_init else None
            o = orc.reset(goal_row(nxt), init_qpos=init)
            assert np.array_equal(o[0], nxt[0]), (prefix, t)
        assert orc.num_interventions[0] == (g["num_interventions"][t] + int(g["reset_after"][t]))
    return g


@pytest.fixture(scope="module")
def rollouts(golden_dir):
    return np.load(os.path.join(golden_dir, "tabletop_ref_rollouts.npz"))


def test_sparse_train_rollout(rollouts):
    g = replay(rollouts, "sparse_train", horizon=1000)
    assert g["done"].sum() == 4 and g["success"].sum() > 100 and (g["obs"][:, 4] == 0).sum() > 50


def test_sparse_eval_rollout(rollouts):
    replay(rollouts, "sparse_eval", horizon=200)


def test_dense_rollout(rollouts):
    replay(rollouts, "dense_train", horizon=700, dense=True)


def test_reset_at_goal_rollout(rollouts):
    replay(rollouts, "resetgoal_train", horizon=300, custom_init=True)


def test_wide_init_rollout(rollouts):
    replay(rollouts, "wide_train", horizon=250, wide=True, custom_init=True)


def test_no_reset_after_done(rollouts):
    g = replay(rollouts, "noreset_train", horizon=100)
    assert g["done"][99:].all() and not g["done"][:99].any() and g["total_steps"][-1] == 300


def test_lifelong(golden_dir):
    g = np.load(os.path.join(golden_dir, "tabletop_ref_lifelong.npz"))
    freq, horizon = int(g["goal_change_frequency"]), int(g["train_horizon"])
    n = len(g["actions"])
    orc = TabletopOracle(1, horizon)
    orc.reset(goal_row(g["obs"][0:1]))
    ll = np.zeros(1)
    since = np.zeros(1, np.int64)
    swap = np.zeros(1, np.uint8)
    for t in range(n):
        ob, rw, dn, _ = orc.step(g["actions"][t][None])
        loader.lib().earl_oracle_lifelong_step(1, rw, ll, since, freq, swap)
        if swap[0]:  # goal swap: new goal drawn by the reference, obs re-read with it (lifelong_wrapper.py:35-42)
            row = goal_row(g["obs"][t + 1:t + 2])
            if not g["reset_after"][t]:
                orc.goal[0, 2:4] = GOAL_STATES[row[0], 2:4]
                ob = orc.get_obs()
        assert rw[0] == g["reward"][t] and dn[0] == g["done"][t] and ll[0] == g["lifelong_return"][t]
        if g["reset_after"][t]:
            # LifelongWrapper.reset -> PersistentStateWrapper.reset; steps_since_goal_change = 0
            since[:] = 0
            ob = orc.reset(goal_row(g["obs"][t + 1:t + 2]))
        assert np.array_equal(ob[0], g["obs"][t + 1]), t


@pytest.mark.parametrize("direction", ["forward", "reverse"])
def test_demo_transitions(golden_dir, direction):
    """Every shipped tabletop transition: oracle == reference step() bit-for-bit; == stored next_obs to 1 ulp."""
    from earl_benchmark_b200 import demos
    d = demos.load("tabletop_manipulation", direction)
    ref = np.load(os.path.join(golden_dir, "tabletop_ref_demo_replay.npz"))
    n = len(d["actions"])
    obs = d["observations"]
    qpos = obs[:, :4].astype(np.float64).copy()
    att = (obs[:, 4] == 0).astype(np.int32)
    goal = obs[:, 6:].astype(np.float64).copy()
    out = np.zeros((n, 12), np.float32)
    rw = np.zeros(n)
    sc = np.zeros(n, np.uint8)
    loader.lib().earl_oracle_tt_step(n, qpos, att, goal, np.ascontiguousarray(d["actions"]), 0, 0, 0, out, rw, sc)
    assert np.array_equal(out, ref[f"{direction}_ref_next_obs"])
    assert np.array_equal(rw, ref[f"{direction}_ref_reward"])
    assert np.abs(out - d["next_observations"]).max() <= 2.4e-7
    assert np.array_equal(out[:, 4:], d["next_observations"][:, 4:])      # attach flags + goal exact
    assert np.array_equal(rw, d["rewards"][:, 0].astype(np.float64))       # 0 reward mismatches
    assert (obs[:, 4] == 0).mean() > 0.4                                   # demos exercise the attach path


def test_fp32_norm_restatement_matches_numpy():
    """np.linalg.norm of an fp32 vector = fp32 products accumulated in index order in fp64, rounded to fp32, fp32
    sqrt: the C restatement against numpy itself on random 2- and 4-vectors (the two sizes the task uses)."""
    import ctypes as C
    from oracle import loader
    L = loader.lib()
    L.earl_oracle_norm_f32.restype = C.c_float
    L.earl_oracle_norm_f32.argtypes = [C.c_void_p, C.c_int]
    rs = np.random.RandomState(5)
    for n in (2, 4):
        for scale in (3.0, 0.15):
            x = rs.uniform(-scale, scale, (4000, n)).astype(np.float32)
            got = np.array([L.earl_oracle_norm_f32(v.ctypes.data, n) for v in x], np.float32)
            assert np.array_equal(got, np.array([np.linalg.norm(v) for v in x], np.float32))
