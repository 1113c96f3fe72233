"""ctypes binding of oracle/libearl_oracle.so (TEST INFRASTRUCTURE ONLY)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_p = np.ctypeslib.ndpointer
_f64 = lambda: _p(np.float64, flags="C_CONTIGUOUS")  # noqa: E731
_f32 = lambda: _p(np.float32, flags="C_CONTIGUOUS")  # noqa: E731
_i64 = lambda: _p(np.int64, flags="C_CONTIGUOUS")  # noqa: E731
_i32 = lambda: _p(np.int32, flags="C_CONTIGUOUS")  # noqa: E731
_u8 = lambda: _p(np.uint8, flags="C_CONTIGUOUS")  # noqa: E731


def build(force=False):
    so = os.path.join(_HERE, "libearl_oracle.so")
    src = os.path.join(_HERE, "tabletop_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libearl_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.earl_oracle_tt_step.argtypes = [C.c_int64, _f64(), _i32(), _f64(), _f32(), C.c_int, C.c_int, C.c_int,
                                          _f32(), _f64(), _u8()]
        L.earl_oracle_tt_reward.argtypes = [C.c_int64, _f32(), C.c_int, C.c_int, _f64(), _u8()]
        L.earl_oracle_psw_step.argtypes = [C.c_int64, _i64(), _i64(), C.c_int64, C.c_void_p, _u8()]
        L.earl_oracle_psw_reset.argtypes = [C.c_int64, C.c_void_p, _i64(), _i64()]
        L.earl_oracle_lifelong_step.argtypes = [C.c_int64, _f64(), _f64(), _i64(), C.c_int64, _u8()]
        roll = [C.c_int64, C.c_int64, C.c_int64, _f64(), _i32(), _f64(), _f32(), C.c_int, C.c_int, C.c_int,
                _i64(), _i64(), C.c_int64, _f32(), _f64(), _u8()]
        L.earl_oracle_tt_rollout.argtypes = roll
        L.earl_oracle_tt_rollout_stepmajor.argtypes = roll
        for f in ("earl_oracle_tt_step", "earl_oracle_tt_reward", "earl_oracle_psw_step", "earl_oracle_psw_reset",
                  "earl_oracle_lifelong_step", "earl_oracle_tt_rollout", "earl_oracle_tt_rollout_stepmajor"):
            getattr(L, f).restype = None
        _LIB = L
    return _LIB


INITIAL_STATE = np.array([0.0, 0.0, 2.5, 0.0, -1.0, -1.0])           # tabletop_manipulation.py:11
GOAL_STATES = np.array([[0.0, 0.0, -2.5, -1.0, -1.0, -1.0],          # tabletop_manipulation.py:12-16
                        [0.0, 0.0, -2.5, 1.0, -1.0, -1.0],
                        [0.0, 0.0, 0.0, 2.0, -1.0, -1.0],
                        [0.0, 0.0, 0.0, -2.0, -1.0, -1.0]])


class TabletopOracle:
    """n independent reference-semantics tabletop envs + PersistentStateWrapper, fp64 state.

    Goal sampling is NOT done here: the caller passes goal rows (the RNG oracle is CPython's own
    `random` module, see tests/test_rng.py).
    """

    def __init__(self, n, horizon, dense=False, wide=False, state_f32=False):
        self.n, self.horizon, self.dense, self.wide, self.state_f32 = n, horizon, int(dense), int(wide), int(state_f32)
        self.qpos = np.zeros((n, 4))
        self.attached = np.zeros(n, np.int32)
        self.goal = np.tile(INITIAL_STATE, (n, 1))
        self.total_steps = np.zeros(n, np.int64)
        self.steps_since_reset = np.zeros(n, np.int64)
        self.num_interventions = np.zeros(n, np.int64)
        self.obs = np.zeros((n, 12), np.float32)
        self.reward = np.zeros(n)
        self.success = np.zeros(n, np.uint8)
        self.done = np.zeros(n, np.uint8)

    def get_obs(self):
        o = np.zeros((self.n, 12), np.float32)
        o[:, :4] = self.qpos.astype(np.float32)
        o[:, 4:6] = np.where(self.attached[:, None] != 0, 0.0, -1.0)
        o[:, 6:] = self.goal.astype(np.float32)
        return o

    def reset(self, goal_rows, mask=None, init_qpos=None):
        """reset() of tabletop_manipulation.py:105-126 (+ wrapper :17-20) for masked envs."""
        m = np.ones(self.n, bool) if mask is None else np.asarray(mask, bool)
        self.attached[m] = 0
        g = np.tile(INITIAL_STATE, (self.n, 1))
        g[:, 2:4] = GOAL_STATES[np.asarray(goal_rows), 2:4]
        self.goal[m] = g[m]
        q = np.tile(INITIAL_STATE[:4], (self.n, 1)) if init_qpos is None else np.asarray(init_qpos, np.float64)
        self.qpos[m] = q[m]
        if self.state_f32:
            self.qpos[m] = self.qpos[m].astype(np.float32).astype(np.float64)
        m8 = np.ascontiguousarray(m, np.uint8)  # keep alive across the call: only its address is passed
        lib().earl_oracle_psw_reset(self.n, m8.ctypes.data, self.steps_since_reset, self.num_interventions)
        return self.get_obs()

    def step(self, action):
        a = np.ascontiguousarray(action, np.float32).reshape(self.n, 3)
        lib().earl_oracle_tt_step(self.n, self.qpos, self.attached, self.goal, a, self.dense, self.wide,
                                  self.state_f32, self.obs, self.reward, self.success)
        lib().earl_oracle_psw_step(self.n, self.total_steps, self.steps_since_reset, self.horizon, None, self.done)
        return self.obs.copy(), self.reward.copy(), self.done.copy(), self.success.copy()
