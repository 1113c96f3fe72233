"""Physics self-checks of the fp64 checker (oracle/mjengine.c) that need no reference data: closed-form free fall of the
free-joint peg, energy conservation of the undamped arm, agreement of the recursive bias forces with a numerical
Lagrangian.  They do not pin the checker to MuJoCo (nothing here can); they pin it to mechanics."""
import copy

import numpy as np
import pytest

from earl_benchmark_b200.envs import sawyer_door, sawyer_peg
from earl_benchmark_b200.mjcf.compile import Model, fk_fused
from oracle.engine import Engine


def _strip(m, damping=True, welds=True, actuators=True, limits=True, collisions=True):
    """Copy of a compiled model with selected mechanisms removed."""
    m = copy.copy(m)
    if welds:
        m.nweld = np.int32(0)
        for k in ("weld_body", "weld_pos", "weld_quat", "weld_relpose", "weld_solref", "weld_solimp", "weld_invweight"):
            setattr(m, k, getattr(m, k)[:0])
    if damping:
        m.dof_damping = np.zeros_like(m.dof_damping)
    if actuators:
        m.act_kp = np.zeros_like(m.act_kp)
    if limits:
        m.jnt_limited = np.zeros_like(m.jnt_limited)
    if collisions:
        m.geom_contype = np.zeros_like(m.geom_contype)
        m.geom_conaffinity = np.zeros_like(m.geom_conaffinity)
    return m


def _energy(m, e):
    M = e.mass_matrix()
    xpos, xmat, _, _ = fk_fused(m, e.qpos)
    com_z = np.array([(xpos[b] + xmat[b] @ m.body_ipos[b])[2] for b in range(int(m.nbody))])
    return 0.5 * e.qvel @ M @ e.qvel + 9.81 * float(np.sum(m.body_mass * com_z))


def test_free_fall_of_the_peg_is_exact():
    """Semi-implicit Euler of a free body under gravity: v_n = -g h n, z_n = z_0 - g h^2 n (n + 1) / 2 (the 0.005 joint
    damping is removed; the huge rotational inertia keeps the orientation fixed)."""
    m = _strip(Model.load(sawyer_peg.MODEL_PATH))
    e = Engine(m)
    e.reset()
    z0, h, n = e.qpos[11], float(m.timestep), 40
    e.step(n)
    assert abs(e.qvel[11] + 9.81 * h * n) < 1e-12
    assert abs(e.qpos[11] - (z0 - 9.81 * h * h * n * (n + 1) / 2)) < 1e-12
    assert np.abs(e.qpos[12:16] - [1, 0, 0, 0]).max() < 1e-12


def test_undamped_arm_conserves_energy():
    """Arm + door swinging violently under gravity (kinetic energy up to 50 J of 72 J) with welds, damping, actuators,
    limits and collisions removed.  Semi-implicit Euler is not symplectic for a configuration-dependent mass matrix, so
    the energy drifts at the task's step size (6 % over 1 s of this swing); the drift must vanish with the step size:
    < 0.5 % at h / 4 and smaller again at h / 16."""
    base = _strip(Model.load(sawyer_door.MODEL_PATH))
    drift = {}
    for div in (1, 4, 16):
        m = copy.copy(base)
        m.timestep = np.float64(float(base.timestep) / div)
        e = Engine(m)
        e.reset()
        e.qpos[:7] = [0.3, -0.9, 0.2, 1.2, -0.4, 0.7, 0.1]
        e0 = _energy(m, e)
        for _ in range(80):
            e.step(5 * div)
        drift[div] = abs(_energy(m, e) - e0) / abs(e0)
    assert drift[4] < 5e-3 and drift[16] < drift[4] < drift[1] < 0.1, drift


def test_bias_forces_match_a_numerical_lagrangian():
    """qfrc_bias = C(q, v) v + g(q) from the recursive Newton-Euler pass against d/dt(dL/dv) - dL/dq evaluated by
    finite differences of the kinetic and potential energy (L = T - V)."""
    m = _strip(Model.load(sawyer_door.MODEL_PATH))
    e = Engine(m)
    rs = np.random.RandomState(1)
    nv = int(m.nv)

    def T_V(q, v):
        e.qpos[:], e.qvel[:] = q, v
        M = e.mass_matrix()
        xpos, xmat, _, _ = fk_fused(m, q)
        com_z = np.array([(xpos[b] + xmat[b] @ m.body_ipos[b])[2] for b in range(int(m.nbody))])
        return 0.5 * v @ M @ v, 9.81 * float(np.sum(m.body_mass * com_z)), M

    for _ in range(3):
        q = rs.uniform(-1, 1, nv) * np.array([1] * 7 + [0.02, 0.02, 0.7])
        v = rs.uniform(-1, 1, nv)
        e.reset()
        e.qpos[:], e.qvel[:] = q, v
        bias = e.bias()
        eps = 1e-6
        # bias_i = sum_j dM_ij/dt v_j - d(T - V)/dq_i  with dM/dt = sum_k dM/dq_k v_k  (all hinge / slide joints)
        dMdt = (T_V(q + eps * v, v)[2] - T_V(q - eps * v, v)[2]) / (2 * eps)
        ref = dMdt @ v
        for i in range(nv):
            dq = np.zeros(nv)
            dq[i] = eps
            Tp, Vp, _ = T_V(q + dq, v)
            Tm, Vm, _ = T_V(q - dq, v)
            ref[i] -= ((Tp - Vp) - (Tm - Vm)) / (2 * eps)
        assert np.abs(bias - ref).max() < 2e-5 * max(1.0, np.abs(ref).max()), (bias, ref)
