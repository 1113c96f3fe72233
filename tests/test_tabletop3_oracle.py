"""The numpy restatement of the three-object tabletop (oracle/tabletop3.py) against outputs of the UNMODIFIED
reference (`earl_benchmark/envs/tabletop_manipulation_3obj.py` under PersistentStateWrapper), committed as
tests/golden/tabletop3_ref_rollouts.npz by oracle/gen_golden_3obj.py.  CPU only."""
import os

import numpy as np
import pytest

from conftest import REPO
from oracle import tabletop3

GOLD = np.load(os.path.join(REPO, "tests", "golden", "tabletop3_ref_rollouts.npz"))


def replay(prefix, horizon, dense, reset_at_goal=False):
    g = {k[len(prefix) + 1:]: GOLD[k] for k in GOLD.files if k.startswith(prefix + "_")}
    n = len(g["actions"])
    orc = tabletop3.Tabletop3Oracle(1, horizon, dense=dense)
    o = orc.reset(init_qpos=g["qpos"][0][None] if reset_at_goal else None)
    assert np.array_equal(o[0], g["obs"][0])
    near = 0
    for t in range(n):
        o, r, d, s = orc.step(g["actions"][t][None])
        if g["reset_after"][t]:
            # what the reference returned from step() is only visible through reward / done / success here
            assert d[0]
        else:
            assert np.array_equal(o[0], g["obs"][t + 1]), t
            assert np.array_equal(orc.qpos[0], g["qpos"][t + 1]), t          # fp64 state, bit for bit
            assert orc.att[0] == g["attached"][t + 1], t
        assert bool(d[0]) == bool(g["done"][t]), t
        assert orc.total_steps == g["total_steps"][t]
        assert orc.num_interventions[0] == g["num_interventions"][t]
        if dense:
            # fixture: fp32 evaluation under numpy 2; restatement: numpy 1.22 promotion (fp64 after the norms)
            assert abs(r[0] - g["reward"][t]) <= 2e-6 * max(1.0, abs(g["reward"][t])), t
        else:
            assert r[0] == g["reward"][t], t
        assert bool(s[0]) == bool(g["success"][t]), t
        if g["reset_after"][t]:
            o = orc.reset(init_qpos=g["qpos"][t + 1][None] if reset_at_goal else None)
            assert np.array_equal(o[0], g["obs"][t + 1])
        near += abs(float(tabletop3.norm_f32(o[0, :8] - o[0, 10:18])) - 0.4) < 1e-5
    return g, near


def test_sparse_rollout_with_horizon_resets():
    g, _ = replay("sparse", 900, dense=False)
    assert g["success"].sum() > 100 and g["done"].sum() == 4 and g["num_interventions"][-1] == 5
    assert set(np.unique(g["attached"])) == {0, 1, 2, 3}


def test_dense_rollout():
    g, _ = replay("dense", 700, dense=True)
    assert g["done"].sum() == 3


def test_reset_at_goal_states_are_taken_from_the_reference_stream():
    g, _ = replay("rag", 50, dense=False, reset_at_goal=True)
    assert g["done"].sum() == 8
    # the reference's reset draws: goal[:8] + np.random.uniform(-0.3, 0.3, 8) after np.random.seed(5)
    rs = np.random.RandomState(5)
    first = tabletop3.GOAL_STATES[0][:8] + rs.uniform(-0.3, 0.3, size=(8,))
    assert np.array_equal(first, g["qpos"][0])


def test_closest_object_attach_cases():
    q0, acts = GOLD["attach_q0"], GOLD["attach_actions"]
    orc = tabletop3.Tabletop3Oracle(len(q0), 1 << 40)
    orc.reset(init_qpos=q0)
    o, _, _, _ = orc.step(acts)
    assert np.array_equal(orc.att, GOLD["attach_att"])
    assert np.array_equal(orc.qpos, GOLD["attach_q1"])
    assert np.array_equal(o, GOLD["attach_obs"])
    assert np.bincount(GOLD["attach_att"], minlength=4).min() > 500


def test_norm_restatement_matches_numpy():
    rs = np.random.RandomState(3)
    for n in (2, 6, 8):
        x = rs.uniform(-3, 3, (5000, n)).astype(np.float32)
        assert np.array_equal(tabletop3.norm_f32(x), np.array([np.linalg.norm(v) for v in x], np.float32))
    d = rs.uniform(-0.35, 0.35, (5000, 2))      # |d|^2 < 0.25: the range where dist_f64 promises the fused form
    assert np.array_equal(tabletop3.dist_f64(d, np.ones(5000, bool)), np.array([np.linalg.norm(v) for v in d]))


def test_constants():
    assert np.array_equal(GOLD["initial_states"][0], tabletop3.INITIAL_STATE)
    assert np.array_equal(GOLD["goal_states"], tabletop3.GOAL_STATES)
