"""mujoco_py stand-in for the reference's KITCHEN task logic -- TEST INFRASTRUCTURE ONLY.

MuJoCo is not in this image, so the reference's kitchen PHYSICS cannot run here.  Everything AROUND the physics is
plain Python / numpy in the reference (KitchenV0.step: action scaling, mocap update and clipping, Robot_VelAct control
from the last noisy observation, do_simulation writing ctrl[0:nu] and calling sim.step() frame_skip times; Robot.get_obs:
observation noise from env.np_random; Kitchen._get_reward_n_score / is_successful; reset_model's np.random draw), and
runs UNMODIFIED against this stand-in, whose `MjSim.step()` advances qpos by a scripted, deterministic pseudo-dynamics.
oracle/gen_kitchen_golden.py records what the reference code computes from those states.

The scripted state evolution and the site positions are INPUTS of the fixtures, not claims about MuJoCo."""
import os

import numpy as np

NQ, NU = 23, 2
SITES = ("microhandle_site", "hinge_site2", "slide_site", "knob1_site", "knob2_site", "knob3_site", "knob4_site",
         "light_site")


class _Opt:
    timestep = 0.002


class PyMjModel:
    def __init__(self, path):
        self.path = path
        self.nq = self.nv = NQ
        self.nu = NU
        self.opt = _Opt()
        self.key_qpos = np.zeros((1, NQ))
        self.key_qvel = np.zeros((1, NQ))
        self.actuator_ctrlrange = np.tile(np.array([[0.0, 0.04]]), (NU, 1))


def load_model_from_path(path):
    if not os.path.isfile(path):
        raise IOError(path)
    return PyMjModel(path)


class _State:
    def __init__(self, qpos, qvel):
        self.qpos, self.qvel = qpos.copy(), qvel.copy()


class _Data:
    def __init__(self, sim):
        self._sim = sim
        self.qpos = np.zeros(NQ)
        self.qvel = np.zeros(NQ)
        self.ctrl = np.zeros(NU)
        self.mocap_pos = np.zeros((1, 3))
        self.mocap_quat = np.array([[1.0, 0, 0, 0]])
        self.time = 0.0

    def get_site_xpos(self, name):
        return self._sim.site_xpos(name)


class MjSim:
    """Scripted pseudo-dynamics: every dof relaxes towards a target that depends on the mocap position and the
    controls, with a dof-dependent rate.  Deterministic, smooth, and exercises every code path of the task logic."""

    def __init__(self, model, nsubsteps=1):
        self.model = model
        self.data = _Data(self)
        self.nsteps = 0
        self.ctrl_log = []

    def reset(self):
        self.data.qpos[:] = 0
        self.data.qvel[:] = 0
        self.data.ctrl[:] = 0
        self.data.mocap_pos[:] = [[0.0, 0.0, 2.89]]
        self.data.time = 0.0

    def forward(self):
        pass

    def step(self):
        d = self.data
        k = np.arange(NQ)
        target = 0.3 * np.sin(1.3 * k + 2.0 * d.mocap_pos[0, 0]) * np.cos(0.7 * k + 3.0 * d.mocap_pos[0, 1]) + 0.1 * d.mocap_pos[0, 2] - 0.2
        target[7:9] = np.clip(d.ctrl[:2], 0.0, 0.04)
        rate = 0.002 * (2.0 + (k % 5))
        new = d.qpos + rate * (target - d.qpos)
        d.qvel[:] = (new - d.qpos) / self.model.opt.timestep
        d.qpos[:] = new
        d.time += self.model.opt.timestep
        self.nsteps += 1
        self.ctrl_log.append(d.ctrl.copy())

    def site_xpos(self, name):
        i = SITES.index(name)
        q = self.data.qpos
        return np.array([-0.6 + 0.15 * i + 0.2 * np.sin(q[9 + i]), 0.4 + 0.1 * np.cos(q[10 + i]), 1.9 + 0.05 * i + 0.1 * q[11 + i]])

    def get_state(self):
        return _State(self.data.qpos, self.data.qvel)

    def set_state(self, s):
        self.data.qpos[:] = s.qpos
        self.data.qvel[:] = s.qvel


class MjViewer:
    def __init__(self, sim):
        raise RuntimeError("no rendering in the stand-in")
