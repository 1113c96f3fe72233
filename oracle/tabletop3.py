"""CPU restatement (numpy, batched) of the three-object tabletop task -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product
package never does.  It follows `earl_benchmark/envs/tabletop_manipulation_3obj.py` (cited per function as
3OBJ:line) under `earl_benchmark/wrappers/persistent_state_wrapper.py` (PSW:line), with a leading env
dimension.  PINNED: tests/test_tabletop3_oracle.py checks it against tests/golden/tabletop3_ref_rollouts.npz,
which oracle/gen_golden_3obj.py produced by running the UNMODIFIED reference class (behind the no-op MuJoCo
stand-in of oracle/fakes; the task is kinematic, `sim.forward()` never changes qpos).

Arithmetic notes (what "the reference computes" means at the bit level, established against numpy here):
  * state and `move` are fp64; np.linalg.norm of the fp64 2-vector = sqrt(fma(dy, dy, dx*dx)) (OpenBLAS ddot);
  * observations are the fp32 casts; np.linalg.norm of an fp32 vector = fp32 products, summed IN ORDER in fp64
    (OpenBLAS sdot accumulates in double), rounded to fp32, fp32 sqrt;
  * `norm <= 0.4` compares the fp32 norm with the fp64 constant (numpy 1.22.2, the reference's pin; numpy 2
    compares in fp32, which differs only if the norm equals float32(0.4) exactly);
  * dense reward: fp32 norms and squares, fp64 from the division by 0.01 onwards (numpy 1.22 scalar promotion).
"""
import math

import numpy as np

INITIAL_STATE = np.array([0.0, 0.0, 2.5, 0.0, 2.5, -1.0, 2.5, 1.0, -1., -1.])      # 3OBJ:11
GOAL_STATES = np.array([[0.0, 0.0, 0.0, -2.0, 0.0, 2.0, -2.5, 1.0, -1., -1.]])      # 3OBJ:12-18
MARKERS = np.array([-1.0, 0.0, 0.5, 1.0])   # attached_object tuples (-1,-1),(0,0),(0.5,0.5),(1,1)  3OBJ:31-37
THRESHOLD, MOVE, CLIP, SUCCESS = 0.4, 0.2, 2.8, 0.4
OBS_DIM = 20


def _fma_scalar(a, b, c):
    """correctly rounded a*b+c (math.fma needs Python >= 3.13: exact rational arithmetic otherwise)"""
    if hasattr(math, "fma"):
        return math.fma(a, b, c)
    from fractions import Fraction
    if not (math.isfinite(a) and math.isfinite(b) and math.isfinite(c)):
        return a * b + c
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def dist_f64(d, need):
    """np.linalg.norm of fp64 2-vectors d[N,2]; rows outside `need` (far from every comparison the caller
    makes) take the unfused sum, which differs by at most one ulp"""
    s = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]
    for i in np.flatnonzero(need & (s < 0.25)):
        s[i] = _fma_scalar(float(d[i, 1]), float(d[i, 1]), float(d[i, 0] * d[i, 0]))
    return np.sqrt(s)


def norm_f32(d):
    """np.linalg.norm over the last axis of an fp32 array, as numpy evaluates it (see module docstring)."""
    d = np.asarray(d, np.float32)
    p = (d * d).astype(np.float64)          # fp32 products, widened
    s = np.zeros(d.shape[:-1], np.float64)
    for k in range(d.shape[-1]):
        s = s + p[..., k]
    return np.sqrt(s.astype(np.float32))


def is_successful(obs):
    """3OBJ:161-165"""
    obs = np.asarray(obs, np.float32)
    return norm_f32(obs[..., :8] - obs[..., 10:18]).astype(np.float64) <= SUCCESS


def compute_reward(obs, dense):
    """3OBJ:146-159 (fp64 result; the device returns its fp32 cast)"""
    obs = np.asarray(obs, np.float32)
    if not dense:
        return is_successful(obs).astype(np.float64)
    r = (-norm_f32(obs[..., 2:8] - obs[..., 12:18])).astype(np.float64)
    for k in range(1, 4):
        nk = norm_f32(obs[..., 2 * k:2 * k + 2] - obs[..., 2 * k + 10:2 * k + 12])
        r = r + 2.0 * np.exp((-(nk * nk)).astype(np.float64) / 0.01)
    return r


class Tabletop3Oracle:
    """N independent reference envs under a PersistentStateWrapper, stepped in lock step."""

    def __init__(self, num_envs, horizon, dense=False, goal_table=GOAL_STATES, initial_state=INITIAL_STATE):
        self.n, self.horizon, self.dense = int(num_envs), int(horizon), bool(dense)
        self.goal_table = np.asarray(goal_table, np.float64)
        self.initial_state = np.asarray(initial_state, np.float64)
        self.qpos = np.zeros((self.n, 8))
        self.att = np.zeros(self.n, np.int64)
        self.goal_row = np.zeros(self.n, np.int64)
        self.steps_since_reset = np.zeros(self.n, np.int64)
        self.num_interventions = np.zeros(self.n, np.int64)
        self.total_steps = 0

    def obs(self):
        """3OBJ:49-54"""
        o = np.empty((self.n, OBS_DIM), np.float32)
        o[:, :8] = self.qpos
        o[:, 8] = o[:, 9] = MARKERS[self.att]
        o[:, 10:] = self.goal_table[self.goal_row]
        return o

    def reset(self, mask=None, goal_rows=None, init_qpos=None):
        """PSW:17-20 + 3OBJ:67-84 (the caller supplies the goal rows and, for reset_at_goal, goal[:8] + noise)"""
        m = np.ones(self.n, bool) if mask is None else np.asarray(mask, bool)
        self.num_interventions[m] += 1
        self.steps_since_reset[m] = 0
        self.att[m] = 0
        self.goal_row[m] = 0 if goal_rows is None else np.asarray(goal_rows)[m]
        self.qpos[m] = self.initial_state[:8] if init_qpos is None else np.asarray(init_qpos, np.float64)[m]
        return self.obs()

    def step(self, actions):
        """PSW:22-31 over 3OBJ:86-144"""
        a = np.clip(np.asarray(actions, np.float32).astype(np.float64), -1.0, 1.0)      # 3OBJ:88
        a = -MOVE + (a + 1.0) * 0.5 * (MOVE - (-MOVE))                                  # 3OBJ:89-90
        fist = self.qpos[:, 0:2].copy()
        grip = a[:, 2] > 0
        free = grip & (self.att == 0)
        held = np.full(self.n, np.inf)
        att = np.where(grip, self.att, 0)                                               # 3OBJ:109-110
        for k in range(3):                                                              # 3OBJ:101-108, dict order
            d = fist - self.qpos[:, 2 + 2 * k:4 + 2 * k]
            dist = dist_f64(d, free)
            take = free & (dist < THRESHOLD) & (dist < held)
            att = np.where(take, k + 1, att)
            held = np.where(take, dist, held)
        nxt = np.clip(fist + a[:, :2], -CLIP, CLIP)                                     # 3OBJ:112-113
        for k in range(3):                                                              # 3OBJ:114-119, 125-127
            sel = att == k + 1
            cur = self.qpos[:, 2 + 2 * k:4 + 2 * k]
            self.qpos[:, 2 + 2 * k:4 + 2 * k] = np.where(sel[:, None], np.clip(cur + (nxt - fist), -CLIP, CLIP), cur)
        self.qpos[:, 0:2] = nxt
        self.att = att
        o = self.obs()
        succ = is_successful(o)
        rew = compute_reward(o, self.dense)
        self.total_steps += 1                                                           # PSW:24-31
        self.steps_since_reset += 1
        done = self.steps_since_reset >= self.horizon
        return o, rew, done, succ
