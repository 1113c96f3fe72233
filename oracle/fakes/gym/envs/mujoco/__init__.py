"""gym.envs.mujoco.MujocoEnv stand-in with a no-op physics backend (test infrastructure only).

Mirrors what gym 0.23.1's `MujocoEnv.__init__` does that is observable to the reference's
tabletop task: build `action_space` from the 3 motor ctrlranges (-1..1,
`tabletop_assets/tabletop_manipulation.xml:95-99`), take ONE random-action `step()` before any
`reset()`, build `observation_space`; `set_state` stores float64 copies and calls `sim.forward()`.
"""
import numpy as np

from gym import Env, spaces


class _Data:
    def __init__(self, nq, nv):
        self.qpos = np.zeros(nq, dtype=np.float64)
        self.qvel = np.zeros(nv, dtype=np.float64)


class _State:
    def __init__(self, time, qpos, qvel):
        self.time, self.qpos, self.qvel, self.act, self.udd_state = time, qpos, qvel, None, {}


class _Model:
    def __init__(self, nq=5, nv=5):
        self.nq, self.nv = nq, nv


class _Sim:
    def __init__(self, nq=5, nv=5):
        self.model = _Model()
        self.data = _Data(nq, nv)
        self.n_forward = 0

    def forward(self):  # mj_forward never writes qpos/qvel
        self.n_forward += 1

    def step(self):
        raise RuntimeError("fake backend has no dynamics; tabletop must never call sim.step()")

    def get_state(self):
        return _State(0.0, self.data.qpos.copy(), self.data.qvel.copy())

    def set_state(self, st):
        self.data.qpos[:] = st.qpos
        self.data.qvel[:] = st.qvel

    def reset(self):
        self.data.qpos[:] = 0.0
        self.data.qvel[:] = 0.0


def _count_dofs(model_path):
    """nq of the tabletop models = their named slide joints outside XML comments (5 for
    tabletop_manipulation.xml, 9 for tabletop_manipulation_3obj.xml)."""
    import re
    try:
        text = open(model_path).read()
    except OSError:
        return 5
    text = re.sub(r"<!--.*?-->", "", text, flags=re.S)
    return len(re.findall(r"<joint\s+name=", text)) or 5


class MujocoEnv(Env):
    def __init__(self, model_path, frame_skip):
        self.frame_skip = frame_skip
        nq = _count_dofs(model_path)
        self.sim = _Sim(nq, nq)
        self.sim.model.nq = self.sim.model.nv = nq
        self.model = self.sim.model
        self.data = self.sim.data
        self.viewer = None
        self.init_qpos = self.sim.data.qpos.ravel().copy()
        self.init_qvel = self.sim.data.qvel.ravel().copy()
        self.action_space = spaces.Box(low=-1.0, high=1.0, shape=(3,), dtype=np.float32)
        action = self.action_space.sample()
        observation, _reward, done, _info = self.step(action)
        assert not done
        self.observation_space = spaces.Box(-np.inf, np.inf, shape=observation.shape, dtype=np.float32)
        self.seed()

    def set_state(self, qpos, qvel):
        assert qpos.shape == (self.model.nq,) and qvel.shape == (self.model.nv,)
        old = self.sim.get_state()
        self.sim.set_state(_State(old.time, np.asarray(qpos, dtype=np.float64), np.asarray(qvel, dtype=np.float64)))
        self.sim.forward()

    @property
    def dt(self):
        return 0.01 * self.frame_skip
