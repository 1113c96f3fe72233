"""Pins the fp64 articulated-body checker (oracle/mjengine.c, oracle/mjcollide.c) against everything the reference ships
for the Sawyer door task: the golden constants of earl_benchmark/envs/sawyer_door.py:13-16 and the ten demonstration
episodes.  MuJoCo itself is not available here, so these are observation-level pins: parity PARTIAL."""
import numpy as np
import pytest

from earl_benchmark_b200 import demos
from earl_benchmark_b200.envs import sawyer_door
from earl_benchmark_b200.mjcf.compile import Model
from oracle.engine import SawyerDoorOracle


def door_angle(handle_xy):
    """Inverse of the handle forward kinematics."""
    p0, hinge = np.array([0.375721629, -0.107139896]), np.array([-0.085, 0.85])
    return np.arctan2(handle_xy[1] - hinge[1], handle_xy[0] - hinge[0]) - np.arctan2(p0[1], p0[0])


@pytest.fixture(scope="module")
def oracle():
    return SawyerDoorOracle(Model.load(sawyer_door.MODEL_PATH))


def test_handle_position_known_answers(oracle):
    """geom 'handle' at door angles -pi/3 and 0 (sawyer_door.py:13-16): MJCF compile incl. legacy mesh centring + FK."""
    for ang, ref in ((-np.pi / 3, sawyer_door.initial_states[0][4:7]), (0.0, sawyer_door.goal_states[0][4:7])):
        ob = oracle.reset(door_angle=ang)
        assert np.abs(ob[4:7] - ref).max() < 5e-8


def test_hand_rest_pose_known_answer(oracle):
    """Hand position after sim.reset() + _reset_hand() (sawyer_door.py:13): a snapshot of a still-moving arm 250
    substeps after a 1.1 m weld pull (the hand is still tilted 29 degrees off the mocap orientation at that instant), so
    it pins weld, limits, bias forces, implicit-damping integration and the constraint regularisers together.  Reached to
    0.65 mm in every coordinate with NO fitted constant (round 1: 1.7 mm with the weld regulariser calibrated to 3.35 x)."""
    ob = oracle.reset()
    assert np.abs(ob[:3] - sawyer_door.initial_states[0][:3]).max() < 1e-3
    assert ob[3] == 1.0


def test_weld_regulariser_is_derived():
    """The weld's translational regulariser is MuJoCo's documented body_invweight0 sum with scale 1.0: the factor round 1
    had to calibrate (2.9 from the rest pose, 3.35 from the door demonstrations) is the inertial frame of the MASSLESS
    `hand` body, which MuJoCo's compiler puts at ipos = pos ("ipos undefined: copy body frame"), 0.12 m off the body origin:
    mj_jacBodyCom there gives 6.106 1/kg instead of 2.168 (x 2.816), the rotational weight is unchanged (284.9)."""
    from earl_benchmark_b200.envs import sawyer_peg
    from earl_benchmark_b200.mjcf import compile as C
    assert C.WELD_TRAN_SCALE == 1.0 and C.MASSLESS_IPOS_FROM_POS
    for path in (sawyer_door.MODEL_PATH, sawyer_peg.MODEL_PATH):
        m = Model.load(path)
        assert abs(m.weld_invweight[0, 0] - 6.1056) < 1e-3 and abs(m.weld_invweight[0, 1] - 284.895) < 1e-2
        assert abs(m.weld_invweight[0, 0] / 2.16812 - 2.816) < 2e-3


def _replay(oracle, task, which):
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import demo_eval
    return demo_eval.replay(oracle, task, which)


def _summary(eps):
    import demo_eval
    return demo_eval.summarise(eps)


def test_forward_door_demonstrations_by_episode(oracle):
    """Five door-closing episodes (75-85 steps, the hand pushing the door shut against the friction of a door panel that
    `obj_init_pos` sinks 23 mm into the table), replayed open loop and judged BY EPISODE: all five reach success; the
    success step is 2-5 steps EARLY (recorded 77, 78, 74, 77, 84; replay 75, 74, 70, 73, 79), i.e. one episode inside the
    +-3 band.  The hand stays within 5.5 cm and the handle within 4 cm of the recording over whole episodes.
    (Round 1's calibrated weld gave 4 of 5 within one step -- by construction: the constant was fitted to these episodes.)"""
    eps = _replay(oracle, "sawyer_door", "forward")
    oracle.goal = oracle.GOAL.copy()
    assert len(eps) == 5 and all(e["success"] for e in eps)
    for e in eps:
        assert -5 <= e["step"] - e["demo_step"] <= 0, (e["step"], e["demo_step"])
        assert e["hand"].max() < 0.055 and e["obj"].max() < 0.04
    s = _summary(eps)
    assert s["within3"] >= 1


def test_sparse_reward_agreement_next_to_the_all_zeros_predictor(oracle, peg_oracle):
    """North-star bar: >= 99 % per-step sparse-reward agreement over the shipped demonstrations -- reported NEXT TO what
    predicting reward 0 everywhere scores, because one success step per episode makes that predictor hard to beat
    (VERDICT r1 weak #1), and by episode.  State of the checker with nothing fitted:
        door forward  5/5 episodes succeed (2-5 steps early)   agreement 0.949   all-zeros 0.987
        door reverse  0/5 (the grasp of the 4 cm handle bar is missed: the free-space approach is up to 16 mm off)
                                                              agreement 0.993 = all-zeros 0.993
        peg  forward  0/10, peg reverse 2/20                   agreement 0.985 / 0.970, all-zeros 0.985 / 0.982
    The bar is NOT met (KNOWN GAP); these assertions keep the numbers honest and fail if they move."""
    rows = {}
    for name, o, task in (("door", oracle, "sawyer_door"), ("peg", peg_oracle, "sawyer_peg")):
        for which in ("forward", "reverse"):
            rows[f"{name}_{which}"] = _summary(_replay(o, task, which))
        o.goal = o.GOAL.copy()
    print({k: (v["success"], v["episodes"], round(v["agreement"], 4), round(v["all_zeros"], 4)) for k, v in rows.items()})
    assert (rows["door_forward"]["success"], rows["door_forward"]["episodes"]) == (5, 5)
    assert rows["door_reverse"]["episodes"] == 5 and rows["peg_forward"]["episodes"] == 10 and rows["peg_reverse"]["episodes"] == 20
    total = sum(v["episodes"] for v in rows.values())
    succ = sum(v["success"] for v in rows.values())
    assert total == 40 and 5 <= succ < 40
    for k, v in rows.items():     # whoever improves the engine must update the documented table above
        assert v["agreement"] >= v["all_zeros"] - 0.04
    assert rows["door_reverse"]["hand_max"] < 0.40


# ------------------------------------------------------------------------------------------------ sawyer_peg
@pytest.fixture(scope="module")
def peg_oracle():
    from earl_benchmark_b200.envs import sawyer_peg
    from oracle.engine import SawyerPegOracle
    return SawyerPegOracle(Model.load(sawyer_peg.MODEL_PATH))


def test_peg_reset_known_answers(peg_oracle):
    """initial_states of sawyer_peg.py:18-50: pegHead = peg position - (0.1, 0, 0) at z = 0.02 (exact: FK of the free
    joint + site), hand rest pose [0.00615235, 0.6001898, 0.19430117] after sim.reset() + _reset_hand().
    KNOWN GAP: the hand rest pose (a snapshot of a moving arm with joint j1 resting ON its limit from substep ~100 on, a
    regime the door scene never visits) is reproduced to 1.1 cm only (x; 1.3 mm in y, 3.3 mm in z)."""
    from earl_benchmark_b200.envs import sawyer_peg
    for row in sawyer_peg.initial_states[:4]:
        ob = peg_oracle.reset(peg_pos=row[4:7] + np.array([0.1, 0, 0]))
        assert np.abs(ob[4:7] - row[4:7]).max() < 1e-6
        assert np.abs(ob[:3] - row[:3]).max() < 0.012
        assert ob[3] == 1.0


def test_gripper_opening_trajectory_known_answer(oracle):
    """obs[3] = clip(|rightEndEffector - leftEndEffector| / 0.1, 0, 1) over the first 40 steps of the first forward door
    demonstration (fingers closing in free space, then running into their soft joint limits: 1.0 -> 0.274 overshoot ->
    0.300).  Independent of the arm, this pins the position actuators (kp 400), armature 100 / damping 1000 with the
    implicit-damping Euler step, the limit rows (solref / solimp / dof_invweight0 regulariser) and the one-substep-stale
    observation timing against the reference's own MuJoCo run: agreement 4e-4 on every step."""
    d = demos.load("sawyer_door", "forward")
    oracle.goal = d["observations"][0][7:14].astype(np.float64)
    oracle.reset(door_angle=door_angle(d["observations"][0][4:6]))
    sim = np.array([oracle.step(d["actions"][t])[0][3] for t in range(40)])
    ref = d["next_observations"][:40, 3]
    oracle.goal = oracle.GOAL.copy()
    assert ref.min() < 0.28 and np.abs(sim - ref).max() < 4e-4, np.abs(sim - ref).max()


def test_recorded_weld_lag_with_the_derived_regulariser(oracle):
    """The mocap position follows from the recorded actions, so hand - mocap is observable in the demonstrations.  While
    the mocap descends at ~0.93 cm per env step (first forward door episode) the RECORDED hand settles 32-34 mm behind it.
    A critically damped weld alone would give 2 * timeconst * v = 28-29 mm (what round 1's checker did with the body-origin
    regulariser); with the derived regulariser (R / A ~ 1 in z at this pose, so joint damping leaks into the lag) the
    checker follows the recorded z-lag within 2 mm on every one of the first nine steps, and the x-lag within 3.5 mm over
    the first five (after that the recorded hand falls up to 9 mm further behind in x: the open gap of DESIGN.md 8.4)."""
    d = demos.load("sawyer_door", "forward")
    obs, nobs, act = d["observations"], d["next_observations"], d["actions"]
    assert np.all(act[:9, 2] < -0.85)                      # steady descent
    o = oracle
    o.goal = obs[0][7:14].astype(np.float64)
    o.reset(door_angle=door_angle(obs[0][4:6]))
    mocap = o.HAND_INIT.copy()
    lag_demo, lag_sim = [], []
    for t in range(9):
        a = np.clip(act[t].astype(np.float64), -1, 1)
        mocap = np.clip(mocap + a[:3] * o.ACTION_SCALE, o.MOCAP_LOW, o.MOCAP_HIGH)
        ob, _ = o.step(act[t])
        lag_demo.append(nobs[t][:3] - mocap)
        lag_sim.append(ob[:3] - mocap)
    o.goal = o.GOAL.copy()
    lag_demo, lag_sim = np.array(lag_demo), np.array(lag_sim)
    assert 0.0315 < lag_demo[6:, 2].max() < 0.0355, lag_demo               # recorded: 32-34 mm
    assert 0.0310 < lag_sim[6:, 2].max() < 0.0340, lag_sim
    assert np.abs(lag_sim[:, 2] - lag_demo[:, 2]).max() < 2e-3, (lag_sim, lag_demo)
    assert np.abs(lag_sim[:5, 0] - lag_demo[:5, 0]).max() < 4.5e-3, (lag_sim, lag_demo)


def test_peg_dense_reward_building_blocks(peg_oracle):
    """metaworld reward_utils as restated for the dense peg reward (sawyer_peg.py:231-299): known values and limits."""
    o = peg_oracle
    assert o._tolerance_long_tail(0.03, 0.0, 0.05, 1.0) == 1.0                      # inside the bounds
    assert abs(o._tolerance_long_tail(1.05, 0.0, 0.05, 1.0) - 0.1) < 1e-12         # value_at_margin at one margin
    assert o._tolerance_long_tail(0.2, 0.0, 0.05, 0.0) == 0.0
    assert o._hamacher(1.0, 0.3) == pytest.approx(0.3) and o._hamacher(0.0, 0.0) == 0.0 and o._hamacher(0.5, 0.5) == pytest.approx(1 / 3)
    zero, one = np.array([0.1, -0.11, 0.01]), np.array([-0.1, -0.15, 0.096])
    assert o._rect_prism_tolerance(np.array([0.3, 0.0, 0.0]), zero, one) == 1.0     # outside the prism
    assert o._rect_prism_tolerance(one.copy(), zero, one) == pytest.approx(1.0)
    assert o._rect_prism_tolerance(0.5 * (zero + one), zero, one) == pytest.approx(0.125)
    ob = o.reset()
    r = o.dense_reward(ob, np.zeros(4))
    assert 0.0 <= r < 1.0                                                           # far from the goal, nothing grasped
    ob2 = ob.copy()
    ob2[4:7] = ob2[11:14]                                                           # peg head at the target
    assert o.dense_reward(ob2, np.zeros(4)) == 10.0
