"""Kitchen on the fp64 checker engine (oracle/engine.py::KitchenOracle).  The task logic is pinned
against the reference (tests/test_kitchen_logic.py); the PHYSICS has no golden data (parity unpinned), so these
are reference-free checks of the engine features the kitchen adds: joint equalities, friction loss, pyramidal cones,
capsule collisions -- and of the task running end to end."""
import os

import numpy as np
import pytest

from earl_benchmark_b200.mjcf.compile import Model
from oracle import kitchen_logic as KL
from oracle.engine import Engine, KitchenOracle

MODEL_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "earl_benchmark_b200", "models", "kitchen.npz")


@pytest.fixture(scope="module")
def model():
    return Model.load(MODEL_PATH)


def test_model_inventory(model):
    """ADEPT/franka/assets/franka_kitchen_jntpos_act_ab.xml as compiled: 23 one-dof joints, 2 finger servos, one mocap
    weld, 5 joint equalities (knob -> burner x174, switch -> light x14), friction loss on every dof, pyramidal cones."""
    assert (int(model.nq), int(model.nv), int(model.nu), int(model.nweld), int(model.neq)) == (23, 23, 2, 1, 5)
    assert int(model.cone_elliptic) == 0 and float(model.timestep) == 0.002 and int(model.iterations) == 100
    assert np.all(model.dof_frictionloss > 0)
    assert model.eq_polycoef[:, 1].tolist() == [174.0] * 4 + [14.0]
    assert model.eq_qposadr.tolist() == [[9, 10], [11, 12], [13, 14], [15, 16], [17, 18]]
    assert sorted(np.unique(model.geom_condim).tolist()) == [3, 6]
    blob = model.to_blob()
    assert np.frombuffer(blob[:8], "<i4").tolist() == [0x4C444D45, 2]        # extended blob: not loadable by the door / peg ABI


def _engine_at(model, q):
    e = Engine(model)
    e.reset()
    e.qpos[:] = q
    e.qvel[:] = 0
    e.mocap_pos[:] = KL.MIDPOINT
    e.forward()
    return e


def test_joint_equality_pulls_the_burner_after_the_knob(model):
    """knob = 174 x burner (polycoef 0 174): start with the knob turned and the burner at rest; the soft equality
    (solref 0.02 1) removes most of the violation within a few time constants, against friction loss on both dofs."""
    q = KL.INIT_QPOS.copy()
    q[9], q[10] = -0.8, 0.0
    e = _engine_at(model, q)
    v0 = abs(e.qpos[9] - 174 * e.qpos[10])
    e.step(200)
    assert abs(e.qpos[9] - 174 * e.qpos[10]) < 0.15 * v0
    assert -0.0091 <= e.qpos[10] <= 1e-6                 # the burner slide stays inside its range


def test_friction_loss_stops_a_coasting_door_and_holds_it(model):
    """The microwave door (hinge, frictionloss 2, damping 2) given 1 rad/s coasts to rest and then does not creep: with
    friction loss alone the force saturates at +-2 N m, so the speed drops at least linearly."""
    q = KL.INIT_QPOS.copy()
    q[22] = -0.7
    e = _engine_at(model, q)
    e.qvel[22] = 1.0
    e.step(5)
    v5 = e.qvel[22]
    assert 0 < v5 < 1.0
    e.step(600)
    assert abs(e.qvel[22]) < 1e-4
    a = e.qpos[22]
    e.step(400)
    assert abs(e.qpos[22] - a) < 1e-5


def test_task_runs_and_contacts_move_the_slide_cabinet(model):
    """Reset + a scripted reach: the arm follows the mocap weld, touches the cabinet (capsule / mesh / box pairs through
    portal refinement, pyramidal rows) and pushes the sliding door open.  Deterministic for fixed seeds."""
    out = []
    for rep in range(2):
        k = KitchenOracle(model)
        k.seed(3)
        np.random.seed(2)
        ob = k.reset()
        assert ob.shape == (46,) and np.array_equal(ob[23:], KL.GOAL)
        e = k.e
        assert np.abs(e.qpos[9:] - KL.ALL_PAIRS[k.config_index, 9:]).max() < 0.02    # objects stay where the reset put them
        target = e.site_xpos("slide_site").copy()
        seen_contact, iters = 0, 0
        for t in range(120):
            d = target - e.site_xpos("end_effector")
            if t >= 70:
                d = np.array([0.5, 0.0, 0.0])
            a = np.zeros(9)
            a[:3] = np.clip(d * 10, -1, 1) * 0.5
            ob, r, s = k.step(a)
            seen_contact = max(seen_contact, e.ncon)
            iters = max(iters, e.solver_iter)
            assert np.all(np.isfinite(ob)) and np.isfinite(r)
        assert seen_contact >= 2 and iters <= 20
        assert e.qpos[19] > 0.1                       # slide cabinet pushed open
        assert r < 0 and s is False
        out.append((e.qpos.copy(), ob.copy(), r))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]) and out[0][2] == out[1][2]


def test_kitchen_known_answers_from_reference_held_constants():
    """What the reference itself holds about the kitchen, pushed through the compiled model: no
    trajectory ships, so these are state-level pins of the MJCF compile + forward kinematics + the (pinned) task logic.
      * `midpoint_pos` (-0.440, 0.1, 2.226), where reset_model parks the mocap (kitchen_multitask_v0.py:46,145-148): the
        end-effector site at INIT_QPOS sits within 2.5 cm of it in x / y (the authors' own hand-picked pair of constants);
      * the goal state (ENV/kitchen.py:28-52) scores exactly 8.0 (eight components done, no distance, no reaching term) and is
        successful; each of the six initial configurations (:57-85) is unsuccessful... unless within 0.3, leaves exactly
        two components to do, scores 6 - 10 ||obj - goal|| - 0.5 ||mocap - site(first unfinished)|| with the site taken from the
        compiled model's kinematics, and that site is the one the reference's table names for that component;
      * site layout the assets fix: the four knob sites form the 0.123 x 0.114 grid of oven_chain.xml."""
    from earl_benchmark_b200.envs import kitchen
    from oracle import kitchen_logic as KL
    from oracle.engine import Engine
    m = Model.load(kitchen.MODEL_PATH)
    e = Engine(m)

    def sites_at(q):
        e.reset()
        e.qpos[:] = q
        e.forward()
        return {s: e.site_xpos(s) for s in m.names["site"]}

    s0 = sites_at(KL.INIT_QPOS)
    assert np.abs(s0["end_effector"][:2] - KL.MIDPOINT[:2]).max() < 0.025
    k = [s0[f"knob{i}_site"] for i in (1, 2, 3, 4)]
    assert np.allclose(k[0] - k[1], [0.123, 0, 0], atol=1e-9) and np.allclose(k[2] - k[0], [0, 0, 0.114], atol=1e-9)
    logic = KL.KitchenLogic()
    goal_obs = np.concatenate([KL.GOAL, KL.GOAL])
    assert logic.reward(goal_obs, KL.MIDPOINT, sites_at(KL.GOAL)) == 8.0 and logic.success(goal_obs)
    for row in KL.ALL_PAIRS:
        q = np.concatenate([KL.INIT_QPOS[:9], row[9:]])
        obs = np.concatenate([q, KL.GOAL])
        todo = [key for key, idx in KL.COMPONENTS if np.linalg.norm(obs[np.array(idx)] - obs[np.array(idx) + 23]) >= len(idx) * 0.01]
        assert len(todo) == 2
        st = sites_at(q)
        expect = 6 - 10 * np.linalg.norm(obs[9:23] - KL.GOAL[9:23]) - 0.5 * np.linalg.norm(KL.MIDPOINT - st[KL.TASK_SITE[todo[0]]])
        assert abs(logic.reward(obs, KL.MIDPOINT, st) - expect) < 1e-12
        assert logic.success(obs) == (np.linalg.norm(obs[9:23] - KL.GOAL[9:23]) <= 0.3)
