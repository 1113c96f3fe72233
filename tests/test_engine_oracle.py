"""Pins the fp64 articulated-body checker (oracle/mjengine.c, oracle/mjcollide.c) against everything the reference ships
for the Sawyer tasks: the golden constants of earl_benchmark/envs/sawyer_door.py:13-16 / sawyer_peg.py:18-50 and the 40
demonstration episodes, all of them OUTPUTS OF THE REFERENCE'S OWN MuJoCo 2.1.0 RUN.  MuJoCo itself is not available here,
so these are observation-level pins -- but tight ones since round 2: wherever no contact between gripper and
object is involved the checker reproduces the recorded fp32 observations to their last digit (hand rest poses to 1e-8 m,
the free-space hand trajectory of all 40 episodes to 1e-7 m, the peg dropped onto the table to 2e-8 m over 26 env steps)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))  # demo_eval

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
    """Hand position after sim.reset() + _reset_hand() (sawyer_door.py:13): a snapshot of a still-moving arm 250 substeps
    after a 1.1 m weld pull (the hand is still tilted 29 degrees off the mocap orientation at that instant), so it pins the
    MJCF compile, kinematics, mass matrix, bias forces, weld, joint limits, finger actuators, the constraint solver and the
    implicit-damping integration together.  Reproduced to 1e-8 m = the fp32 rounding of the reference constant itself, with
    NO fitted constant (round 1: 1.7 mm with the weld regulariser calibrated to 3.35 x)."""
    ob = oracle.reset()
    assert np.abs(ob[:3] - sawyer_door.initial_states[0][:3]).max() < 2e-8
    assert ob[3] == 1.0


def test_weld_regulariser_is_derived():
    """Two facts about MuJoCo 2.1.0 established against the reference's own data, no calibration left:
    (i) the inertial frame of a MASSLESS body sits at ipos = pos ("ipos undefined: copy body frame"), 0.12 m off the origin of
    the `hand` body: mj_jacBodyCom there gives a translational body_invweight0 of 6.106 1/kg instead of 2.168 (x 2.816: what
    round 1 had fitted as 2.9 - 3.35); (ii) mj_diagApprox gives ALL SIX weld rows that translational weight (the rotational
    weight, 284.9, is 46.7 x larger: with it the rest poses are off by 0.65 mm / 10 mm and the free-space tracking by 7 mm)."""
    from earl_benchmark_b200.envs import sawyer_peg
    from earl_benchmark_b200.mjcf import compile as C
    assert C.WELD_TRAN_SCALE == 1.0 and C.MASSLESS_IPOS_FROM_POS and C.WELD_ROT_USES_TRAN_WEIGHT
    for path in (sawyer_door.MODEL_PATH, sawyer_peg.MODEL_PATH):
        m = Model.load(path)
        assert abs(m.weld_invweight[0, 0] - 6.1056) < 1e-3 and m.weld_invweight[0, 1] == m.weld_invweight[0, 0]
        assert abs(m.weld_invweight[0, 0] / 2.16812 - 2.816) < 2e-3


def test_free_space_hand_trajectories_of_all_forty_demonstrations(oracle, peg_oracle):
    """Open-loop replay of the recorded actions from the reconstructed start state, up to two steps before the first
    contact (door: until the handle first moves; peg: 8 steps): 147 door + 240 peg transitions.  Hand position and gripper
    opening agree with the recording to 1e-7 m / 1e-7 -- the precision of the stored fp32 values -- on every one of them."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import freespace_fit
    r = freespace_fit.metric(oracle.e.model, peg_oracle.e.model)
    assert r["sawyer_door_n"] == 147 and r["sawyer_peg_n"] == 240
    for task in ("sawyer_door", "sawyer_peg"):
        assert np.abs(r[task + "_rest_mm"]).max() < 2e-5          # mm
        assert r[task + "_max_mm"].max() < 2e-4, r                 # mm
        assert r[task + "_grip_max"] < 1e-7


def test_peg_dropped_on_the_table_matches_the_recording(peg_oracle):
    """Every peg episode starts with the peg released 5 mm above the table (`_set_obj_xyz` writes z = 0.02, the peg's half
    thickness is 0.015): free fall, a box-box face contact with the table, 2.3 mm of penetration, an overdamped creep back
    to a 0.22 mm rest penetration over ~25 env steps.  With MuJoCo 2.1.0's contact distance for box-box face contacts (HALF the
    vertex-below-face depth) the checker follows the recorded pegHead height to 2e-8 m on all of the first 26 steps (0.87 mm
    off with the full depth): this pins the contact rows (solref / solimp mixing, impedance, regulariser from
    body_invweight0, reference acceleration) and the free joint against the reference's own MuJoCo run."""
    d = demos.load("sawyer_peg", "forward")
    obs, nobs, act = d["observations"], d["next_observations"], d["actions"]
    peg_oracle.goal = obs[0][7:14].astype(np.float64)
    peg_oracle.reset(peg_pos=demos.peg_position_from_obs(obs[0]).astype(np.float64))
    err = [np.abs(peg_oracle.step(act[t])[0][4:7] - nobs[t][4:7]).max() for t in range(26)]
    peg_oracle.goal = peg_oracle.GOAL.copy()
    assert nobs[3][6] < 0.0127 < nobs[25][6] < 0.0148          # the landing is in the data
    assert max(err) < 5e-8, max(err)


def _replay(oracle, task, which):
    import demo_eval
    return demo_eval.replay(oracle, task, which)


def _summary(eps):
    import demo_eval
    return demo_eval.summarise(eps)


def test_demonstrations_by_episode_next_to_the_all_zeros_predictor(oracle, peg_oracle):
    """North-star bar: >= 99 % per-step sparse-reward agreement when the shipped demonstrations are replayed -- judged BY
    EPISODE and reported next to what predicting reward 0 everywhere scores (one success step per episode makes that
    predictor hard to beat).  State of the fp64 checker with nothing fitted:

        set           reaching success (+ up to 3 held steps)   success step within +-3   per-step agreement   all-zeros
        peg forward           6 (10) / 10                               10                      0.9898           0.9854
        peg reverse          13 (19) / 20                               19                      0.9938           0.9823
        door forward          5      /  5                                5  (1-3 steps early)   0.9671           0.9873
        door reverse          5      /  5                                5  (0-1 steps early)   0.9957           0.9929

    39 OF THE 40 EPISODES end within +-3 steps of the recording.  The recordings end ON their success step (peg: 47-49.8 mm from
    the goal, radius 50 mm), so a replay that is one step late is "unsuccessful" inside the recording: "+3" is judged by holding
    the last recorded action for up to three more steps (tools/demo_eval.py).  PEG: 0.9923 over the 1,815 peg transitions (bar:
    0.99), above the null predictor; the one episode that stays out starts with the peg inside the block walls.  DOOR: free space
    and first contact are exact (1e-5 rad after the first contact step); the gripper grasps the handle (mjc_fixNormal) and the
    door follows the recording to 2.2 cm over whole episodes.  The forward set's per-step agreement (0.967) stays below the null
    predictor's 0.987: 13 early success steps in 395 transitions."""
    import demo_eval
    rows = {}
    for name, o, task in (("door", oracle, "sawyer_door"), ("peg", peg_oracle, "sawyer_peg")):
        for which in ("forward", "reverse"):
            rows[f"{name}_{which}"] = demo_eval.summarise(_replay(o, task, which))
        o.goal = o.GOAL.copy()
    print({k: (v["success"], v["episodes"], v["within3"], round(v["agreement"], 4), round(v["all_zeros"], 4)) for k, v in rows.items()})
    pf, pr, df, dr = rows["peg_forward"], rows["peg_reverse"], rows["door_forward"], rows["door_reverse"]
    assert (pf["episodes"], pr["episodes"], df["episodes"], dr["episodes"]) == (10, 20, 5, 5)
    assert pf["success_incl_3_more_steps"] == 10 and pf["within3"] == 10 and pf["agreement"] > 0.989 > pf["all_zeros"]
    assert pr["success"] >= 13 and pr["within3"] >= 19 and pr["agreement"] > 0.99 > pr["all_zeros"]
    assert pf["within3"] + pr["within3"] >= 29                            # VERDICT r1 bar for the peg: >= 27 of 30 within +-3
    peg_total = (pf["agreement"] * 683 + pr["agreement"] * 1132) / 1815
    assert peg_total >= 0.99, peg_total                                   # north-star bar, peg task
    assert pf["hand_max"] < 0.02 and pr["hand_max"] < 0.008               # the hand stays within 2 cm / 8 mm for whole episodes
    assert df["success"] == 5 and df["within3"] == 5 and df["hand_max"] < 0.025 and df["obj_max"] < 0.022
    assert dr["success"] == 5 and dr["within3"] == 5 and dr["obj_max"] < 0.025
    assert dr["agreement"] > 0.995 > dr["all_zeros"]
    assert df["agreement"] > 0.965                                         # 13 early steps in 395; the null predictor has 0.987


def test_solref_mixing_rule_is_decided_by_the_recordings(oracle):
    """MODEL SELECTION among parameter-free candidates, not a fit.  Every mixed contact pair of these scenes combines the time
    constants 0.02 (default geoms: table, peg, claws) and 0.01 (door / block collision geoms, gripper pads).  MuJoCo's
    documentation says the pair gets a "weighted average"; averaging the time constants themselves (0.015) leaves the door panel,
    which drags over the table it is sunk into, with 11 % too little friction damping (B = 2 / (dmax tc)) and a 21 % smaller
    friction cone (K = 1 / (dmax tc)^2 on a normal row whose Jacobian is zero): every door episode then ends 4-12 steps early
    and the gripper loses the handle in one.  Averaging the INVERSE time constants (harmonic mean, 0.01333: the geoms'
    natural frequencies) puts all ten door episodes within +-3 steps.  The peg episodes prefer neither strongly (28 vs 29 of 30
    within +-3).  This test keeps the evidence: the door under both rules."""
    import demo_eval
    from oracle.engine import lib
    L = lib()
    try:
        L.mje_set_opt(8, 1.0)      # arithmetic mean of the time constants
        arith = {w: demo_eval.summarise(_replay(oracle, "sawyer_door", w)) for w in ("forward", "reverse")}
    finally:
        L.mje_set_opt(8, 0.0)
        oracle.goal = oracle.GOAL.copy()
    harm = {w: demo_eval.summarise(_replay(oracle, "sawyer_door", w)) for w in ("forward", "reverse")}
    oracle.goal = oracle.GOAL.copy()
    assert arith["forward"]["within3"] == 0 and arith["reverse"]["within3"] == 0 and arith["reverse"]["success"] == 4
    assert harm["forward"]["within3"] == 5 and harm["reverse"]["within3"] == 5
    assert arith["forward"]["obj_max"] > 0.04 > 0.022 > harm["forward"]["obj_max"]      # handle tracking error over whole episodes
    assert harm["forward"]["agreement"] > arith["forward"]["agreement"] + 0.05
    assert harm["reverse"]["agreement"] > arith["reverse"]["agreement"] + 0.03


# ------------------------------------------------------------------------------------------------ sawyer_peg
@pytest.fixture(scope="module")
def peg_oracle():
    from earl_benchmark_b200.envs import sawyer_peg
    from oracle.engine import SawyerPegOracle
    return SawyerPegOracle(Model.load(sawyer_peg.MODEL_PATH))


def test_peg_reset_known_answers(peg_oracle):
    """initial_states of sawyer_peg.py:18-50: pegHead = peg position - (0.1, 0, 0) at z = 0.02 (exact: FK of the free
    joint + site), hand rest pose [0.00615235, 0.6001898, 0.19430117] after sim.reset() + _reset_hand().
    The rest pose (a snapshot of a moving arm with joint j1 resting ON its limit from substep ~100 on, a regime the door scene
    never visits; round 1: 1.2 cm off) is reproduced to 1e-8 m."""
    from earl_benchmark_b200.envs import sawyer_peg
    for row in sawyer_peg.initial_states[:4]:
        ob = peg_oracle.reset(peg_pos=row[4:7] + np.array([0.1, 0, 0]))
        assert np.abs(ob[4:7] - row[4:7]).max() < 1e-6
        assert np.abs(ob[:3] - row[:3]).max() < 2e-8
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


def test_recorded_weld_lag(oracle):
    """The mocap position follows from the recorded actions, so hand - mocap is observable in the demonstrations.  While the
    mocap descends at ~0.93 cm per env step (first forward door episode) the RECORDED hand settles 32-34 mm behind it -- not
    the 28-29 mm of a critically damped weld alone (2 * timeconst * v; round 1's checker), because joint damping leaks into
    the lag through the weld's regulariser.  The checker reproduces the recorded lag in x, y and z to 1e-7 m on each of the
    first nine steps."""
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
    assert np.abs(lag_sim - lag_demo).max() < 2e-7, (lag_sim, lag_demo)


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
