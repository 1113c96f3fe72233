"""Kitchen task logic around the physics (SURVEY 8 rows a12 glue, a13, a14, reset draw): the numpy restatement against
what the reference's own code computed behind a scripted MuJoCo stand-in (oracle/gen_kitchen_golden.py)."""
import os

import numpy as np
import pytest

from oracle import kitchen_logic as KL


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "kitchen_ref_logic.npz"))


def test_constants_match_the_reference(gold):
    assert np.array_equal(gold["goal_states"][0], KL.GOAL)
    assert np.array_equal(gold["all_pairs"], KL.ALL_PAIRS)
    assert np.array_equal(gold["init_qpos"], KL.INIT_QPOS)
    assert np.array_equal(gold["pos_noise_amp"][:23], KL.POS_NOISE_AMP)
    assert np.array_equal(gold["pos_bound"][:23], KL.POS_BOUND)
    assert np.array_equal(gold["vel_bound"][:23], KL.VEL_BOUND)
    assert np.array_equal(gold["midpoint_pos"], KL.MIDPOINT)
    assert np.array_equal(gold["mocap_clip"], np.stack([KL.MOCAP_LOW, KL.MOCAP_HIGH]))
    assert int(gold["frame_skip"]) == KL.FRAME_SKIP and float(gold["noise_ratio"]) == KL.NOISE_RATIO
    assert tuple(gold["site_names"]) == KL.SITES


def test_step_glue_observation_reward_bit_exact(gold):
    """Replays every recorded episode: the stream consumption of reset (6 observations) and step (1), the mocap update,
    the two controls computed from the LAST NOISY observation, the noisy observation and the reward must be identical
    to the last bit."""
    for e in range(int(gold["n_episodes"])):
        g = lambda k: gold[f"ep{e}_{k}"]  # noqa: E731
        env_seed, np_seed = (int(x) for x in g("seeds"))
        k = KL.KitchenLogic()
        k.seed(env_seed)
        np.random.seed(np_seed)
        q0, idx = k.reset_state()
        # robot.reset refreshes the observation cache (5 observations of the reset pose), then 10 settle steps whose
        # controls come from the last of them, then the observation the reset returns
        for _ in range(5):
            k.observe(q0, noise_ratio=1)                            # get_obs default ratio: 10 x the observation noise
        mocap, ctrl = k.control(np.zeros(9), KL.MIDPOINT)           # reset_model drives robot.step directly: no mocap move
        assert np.array_equal(ctrl, g("reset_ctrl"))
        assert int(g("reset_sim_steps")) == 10 * KL.FRAME_SKIP
        ob = k.observe(g("reset_qpos"))
        assert np.array_equal(ob, g("reset_obs"))
        assert np.array_equal(g("reset_mocap"), KL.MIDPOINT)
        assert k.success(ob) == bool(g("reset_success"))
        assert np.random.randint(1 << 30) == int(g("reset_next_np_random"))   # exactly one legacy draw per reset
        mocap = KL.MIDPOINT.copy()
        for t, a in enumerate(g("actions")):
            assert np.array_equal(mocap, g("mocap_before")[t])
            mocap, ctrl = k.control(a, mocap)
            assert np.array_equal(mocap, g("mocap")[t])
            assert np.array_equal(ctrl, g("ctrl")[t])
            assert int(g("sim_steps")[t]) == KL.FRAME_SKIP
            ob = k.observe(g("qpos")[t])
            assert np.array_equal(ob, g("obs")[t])
            assert k.reward(ob, mocap, g("sites")[t]) == g("reward")[t]
            assert k.success(ob) == bool(g("success")[t])


def test_reset_configuration_follows_the_drawn_index(gold):
    for e in range(int(gold["n_episodes"])):
        np.random.seed(int(gold[f"ep{e}_seeds"][1]))
        q0, idx = KL.KitchenLogic.reset_state()
        # the scripted stand-in relaxes qpos, so compare what the reset WROTE: objects of the drawn configuration
        assert 0 <= idx < 6 and np.array_equal(q0[9:], KL.ALL_PAIRS[idx, 9:]) and np.array_equal(q0[:9], KL.INIT_QPOS[:9])


def test_goal_reaching_branches(gold):
    k = KL.KitchenLogic()
    ob = gold["near_goal_obs"]
    r = k.reward(ob, gold["near_goal_mocap"], gold["near_goal_sites"])
    assert r == float(gold["near_goal_reward"])
    assert k.success(ob) is True and bool(gold["near_goal_success"]) is True
    assert r > 0          # several components inside their 0.01-per-index band: the +1 branch fired


def test_product_constants_match_the_reference(gold):
    """envs/kitchen.py carries its own copy of the task constants (the product never imports the checker): same values as
    the reference's objects at run time."""
    from earl_benchmark_b200.envs import kitchen as kt
    assert np.array_equal(kt.goal_states, gold["goal_states"])
    assert np.array_equal(kt.initial_states["all_pairs"], gold["all_pairs"])
    assert np.array_equal(kt.INIT_QPOS, gold["init_qpos"])
    assert np.array_equal(kt.POS_NOISE_AMP, gold["pos_noise_amp"][:23])
    assert np.array_equal(kt.POS_BOUND, gold["pos_bound"][:9]) and np.array_equal(kt.VEL_BOUND, gold["vel_bound"][:9])
    assert np.array_equal(kt.MIDPOINT, gold["midpoint_pos"])
    assert np.array_equal(np.stack([kt.MOCAP_LOW, kt.MOCAP_HIGH]), gold["mocap_clip"])
    assert kt.FRAME_SKIP == int(gold["frame_skip"]) and kt.NOISE_RATIO == float(gold["noise_ratio"])
    # reward sites in component order (kitchen.py:15-25 with :149-156)
    assert kt.REWARD_SITES == tuple(KL.TASK_SITE[name] for name, _ in KL.COMPONENTS)
