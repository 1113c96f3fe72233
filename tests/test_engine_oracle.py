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
    substeps after a 1 m weld pull, so it pins weld, limits, bias forces and integration together.  Reached to 2 mm."""
    ob = oracle.reset()
    assert np.abs(ob[:3] - sawyer_door.initial_states[0][:3]).max() < 2e-3
    assert ob[3] == 1.0


def _replay(oracle, which):
    d = demos.load("sawyer_door", which)
    obs, nobs, act = d["observations"], d["next_observations"], d["actions"]
    term, rew = d["terminals"].ravel(), d["rewards"].ravel()
    ends = list(np.nonzero(term)[0] + 1)
    starts = [0] + ends[:-1]
    out = []
    for s, en in zip(starts, ends):
        oracle.goal = obs[s][7:14].astype(np.float64)
        oracle.reset(door_angle=door_angle(obs[s][4:6]))
        r, hand, handle = [], [], []
        for t in range(s, en):
            ob, rr = oracle.step(act[t])
            r.append(rr)
            hand.append(np.abs(ob[:3] - nobs[t][:3]).max())
            handle.append(np.abs(ob[4:7] - nobs[t][4:7]).max())
        out.append(dict(reward=np.array(r), demo_reward=rew[s:en], hand=np.array(hand), handle=np.array(handle)))
    oracle.goal = oracle.GOAL.copy()
    return out


def test_forward_demonstrations_replay_open_loop(oracle):
    """Five door-closing episodes (75-85 steps each, hand pushing the door against table friction): the open-loop
    replay keeps the hand within 1.5 cm and the handle within 1.2 cm of the recorded trajectory for the whole episode;
    four of the five episodes close the door within one step of the recorded success step, the fifth is 1 cm short
    of the goal when the recording ends."""
    eps = _replay(oracle, "forward")
    assert len(eps) == 5
    exact = 0
    for ep in eps:
        assert ep["hand"].max() < 0.015 and ep["handle"].max() < 0.012
        demo_step = int(np.nonzero(ep["demo_reward"])[0][0])
        first = np.nonzero(ep["reward"])[0]
        exact += int(len(first) > 0 and abs(int(first[0]) - demo_step) <= 1)
    assert exact >= 4


def test_sparse_reward_agreement_over_all_demonstrations(oracle):
    """North-star bar: >= 99 % per-step sparse-reward agreement over the shipped demonstrations (1,095 transitions).
    The five reverse (grasp-and-pull) episodes track the recording until the grasp and then lose the handle, so
    their single success step each is counted as a mismatch."""
    eps = _replay(oracle, "forward") + _replay(oracle, "reverse")
    total = sum(len(e["reward"]) for e in eps)
    mism = sum(int((e["reward"] != e["demo_reward"]).sum()) for e in eps)
    assert total == 1095
    assert 1 - mism / total >= 0.99, (mism, total)
    for ep in eps[5:]:  # reverse episodes: free-space approach (first 45 steps) tracks the recording
        assert ep["hand"][:45].max() < 0.03


# ------------------------------------------------------------------------------------------------ sawyer_peg
@pytest.fixture(scope="module")
def peg_oracle():
    from earl_benchmark_b200.envs import sawyer_peg
    from oracle.engine import SawyerPegOracle
    return SawyerPegOracle(Model.load(sawyer_peg.MODEL_PATH))


def test_peg_reset_known_answers(peg_oracle):
    """initial_states of sawyer_peg.py:18-50: pegHead = peg position - (0.1, 0, 0) at z = 0.02 (exact: FK of the free
    joint + site), hand rest pose [0.00615235, 0.6001898, 0.19430117] after sim.reset() + _reset_hand().
    KNOWN GAP: the hand rest pose (a snapshot of a moving arm with joint j1 on its limit) is reproduced to 1.2 cm only."""
    from earl_benchmark_b200.envs import sawyer_peg
    for row in sawyer_peg.initial_states[:4]:
        ob = peg_oracle.reset(peg_pos=row[4:7] + np.array([0.1, 0, 0]))
        assert np.abs(ob[4:7] - row[4:7]).max() < 1e-6
        assert np.abs(ob[:3] - row[:3]).max() < 0.012
        assert ob[3] == 1.0


def test_peg_demonstrations_are_not_reproduced(peg_oracle):
    """Documents the state of the peg checker honestly: the open-loop replay of the shipped forward demonstrations
    tracks the recorded HAND within 7 cm, but the grasp-lift-insert sequence is not reproduced (the peg is left behind),
    so no episode reaches the goal and the per-step sparse agreement (98.4 %) is below the 99 % north-star bar."""
    d = demos.load("sawyer_peg", "forward")
    obs, nobs, act = d["observations"], d["next_observations"], d["actions"]
    term, rew = d["terminals"].ravel(), d["rewards"].ravel()
    ends = list(np.nonzero(term)[0] + 1)
    total = mism = 0
    for s, en in zip([0] + ends[:-1], ends):
        peg_oracle.goal = obs[s][7:14].astype(np.float64)
        peg_oracle.reset(peg_pos=obs[s][4:7].astype(np.float64) + np.array([0.1, 0, 0]))
        r, hand = [], []
        for t in range(s, en):
            ob, rr = peg_oracle.step(act[t])
            r.append(rr)
            hand.append(np.abs(ob[:3] - nobs[t][:3]).max())
        assert max(hand) < 0.07
        total += en - s
        mism += int((np.array(r) != rew[s:en]).sum())
    peg_oracle.goal = peg_oracle.GOAL.copy()
    assert 0.98 <= 1 - mism / total < 0.99


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


def test_recorded_weld_lag_vs_documented_and_calibrated_weld():
    """KNOWN GAP, kept measurable.  The mocap position follows from the recorded actions, so hand - mocap is observable
    in the demonstrations.  While the mocap descends at ~0.93 cm per env step the RECORDED hand settles 32-34 mm behind
    it.  With MuJoCo's documented weld regulariser the checker settles at the critically damped value
    2 * timeconst * v (minus one substep of observation staleness) = 28-29 mm; with the calibrated translational
    regulariser (`WELD_TRAN_SCALE` = 3.35) it follows the recording within 2 mm."""
    from earl_benchmark_b200.mjcf.compile import WELD_TRAN_SCALE
    d = demos.load("sawyer_door", "forward")
    obs, nobs, act = d["observations"], d["next_observations"], d["actions"]
    assert np.all(act[:9, 2] < -0.85)                      # steady descent
    lags = {}
    for name, scale in (("calibrated", 1.0), ("documented", 1.0 / WELD_TRAN_SCALE)):
        m = Model.load(sawyer_door.MODEL_PATH)
        m.weld_invweight[:, 0] *= scale
        o = SawyerDoorOracle(m)
        o.goal = obs[0][7:14].astype(np.float64)
        o.reset(door_angle=door_angle(obs[0][4:6]))
        mocap = o.HAND_INIT.copy()
        lag_demo, lag_sim = [], []
        for t in range(9):
            a = np.clip(act[t].astype(np.float64), -1, 1)
            mocap = np.clip(mocap + a[:3] * o.ACTION_SCALE, o.MOCAP_LOW, o.MOCAP_HIGH)
            ob, _ = o.step(act[t])
            lag_demo.append(nobs[t][2] - mocap[2])
            lag_sim.append(ob[2] - mocap[2])
        lags[name] = np.array(lag_sim)
    lag_demo = np.array(lag_demo)
    assert 0.0315 < lag_demo[6:].max() < 0.0355, lag_demo               # recorded: 32-34 mm
    assert 0.0275 < lags["documented"][6:].max() < 0.0300, lags         # 2 * 0.02 s * 0.74 m/s - staleness
    assert np.abs(lags["calibrated"][3:] - lag_demo[3:]).max() < 2e-3, (lags, lag_demo)


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
