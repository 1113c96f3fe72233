"""ctypes binding of oracle/libearl_mjengine.so + the metaworld / EARL Sawyer env logic on top of it.
TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
PTR = dict(qpos=0, qvel=1, ctrl=2, mocap_pos=3, mocap_quat=4, xpos=5, site_xpos=6, geom_xpos=7, qacc=8, qacc_warmstart=9, M=10,
           qfrc_bias=11, efc_force=12, efc_pos=13, qfrc_constraint=14, xmat=15, qfrc_smooth=16, qacc_smooth=17, con_pos=18,
           con_dist=19, con_frame=20)


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libearl_mjengine.so")
        srcs = [os.path.join(_HERE, f) for f in ("mjengine.c", "mjcollide.c", "mjengine.h")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["make", "-C", _HERE, "-B", "libearl_mjengine.so"], stdout=subprocess.DEVNULL)
        L = C.CDLL(so)
        L.mje_load.restype = C.c_void_p
        L.mje_load.argtypes = [C.c_char_p, C.c_longlong]
        L.mje_make_data.restype = C.c_void_p
        L.mje_ptr.restype = C.POINTER(C.c_double)
        L.mje_ptr.argtypes = [C.c_void_p, C.c_int]
        L.mje_int.argtypes = [C.c_void_p, C.c_int]
        L.mje_flops.restype = C.c_longlong
        L.mje_flops.argtypes = [C.c_void_p]
        for f in ("mje_reset", "mje_forward", "mje_step", "mje_kinematics", "mje_mass_matrix", "mje_bias"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_void_p]
            getattr(L, f).restype = None
        L.mje_multi_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.mje_free.argtypes = [C.c_void_p]
        L.mje_set_opt.argtypes = [C.c_int, C.c_double]
        L.mje_debug_boxbox_face_scale.argtypes = [C.c_double]
        L.mje_con_geoms.argtypes = [C.c_void_p, C.c_int]
        L.mje_debug_fix_normal.argtypes = [C.c_int]
        L.mje_free_data.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


class Engine:
    """One fp64 environment instance of the fused model `model` (earl_benchmark_b200.mjcf.compile.Model)."""

    def __init__(self, model):
        self.model = model
        blob = model.to_blob()
        self.m = lib().mje_load(blob, len(blob))
        if not self.m:
            raise RuntimeError("mje_load rejected the model blob")
        self.d = lib().mje_make_data()
        self.nq, self.nv = int(model.nq), int(model.nv)
        self.reset()

    def __del__(self):
        try:
            lib().mje_free_data(self.d)
            lib().mje_free(self.m)
        except Exception:
            pass

    def arr(self, name, shape):
        p = lib().mje_ptr(self.d, PTR[name])
        return np.ctypeslib.as_array(p, shape=(int(np.prod(shape)),)).reshape(shape)

    qpos = property(lambda s: s.arr("qpos", (s.nq,)))
    qvel = property(lambda s: s.arr("qvel", (s.nv,)))
    ctrl = property(lambda s: s.arr("ctrl", (int(s.model.nu),)))
    mocap_pos = property(lambda s: s.arr("mocap_pos", (3,)))
    mocap_quat = property(lambda s: s.arr("mocap_quat", (4,)))
    qacc = property(lambda s: s.arr("qacc", (s.nv,)))
    nefc = property(lambda s: lib().mje_int(s.d, 0))
    ncon = property(lambda s: lib().mje_int(s.d, 1))
    solver_iter = property(lambda s: lib().mje_int(s.d, 2))
    flops = property(lambda s: lib().mje_flops(s.d))

    def contacts(self):
        """[(geom name 1, geom name 2, dist)] of the last forward pass."""
        names, dist = self.model.names["geom"], self.arr("con_dist", (64,))
        out = []
        for k in range(self.ncon):
            c = lib().mje_con_geoms(self.d, k)
            out.append((names[c // 1000], names[c % 1000], float(dist[k])))
        return out

    def reset(self):
        lib().mje_reset(self.m, self.d)

    def forward(self):
        lib().mje_forward(self.m, self.d)

    def step(self, n=1):
        lib().mje_multi_step(self.m, self.d, int(n))

    def site_xpos(self, name):
        return self.arr("site_xpos", (40, 3))[self.model.site_id(name)].copy()

    def geom_xpos(self, name):
        return self.arr("geom_xpos", (160, 3))[self.model.geom_id(name)].copy()

    def mass_matrix(self):
        lib().mje_kinematics(self.m, self.d)
        lib().mje_mass_matrix(self.m, self.d)
        return self.arr("M", (32, 32))[:self.nv, :self.nv].copy()

    def bias(self):
        lib().mje_kinematics(self.m, self.d)
        lib().mje_bias(self.m, self.d)
        return self.arr("qfrc_bias", (32,))[:self.nv].copy()


class SawyerDoorOracle:
    """metaworld SawyerXYZEnv.step / reset semantics + EARL SawyerDoorV2 observation and sparse reward
    (reference earl_benchmark/envs/sawyer_door.py:86-125,141-177)."""
    MOCAP_LOW = np.array([-0.5, 0.40, 0.05])
    MOCAP_HIGH = np.array([0.5, 1.0, 0.5])
    HAND_INIT = np.array([0, 0.4, 0.2], dtype=np.float32).astype(np.float64)
    GOAL = np.array([0.29072163, 0.74286009, 0.10003595, 1.0, 0.29072163, 0.74286009, 0.10003595])
    FRAME_SKIP, ACTION_SCALE = 5, 1.0 / 100

    def __init__(self, model):
        self.e = Engine(model)
        self.goal = self.GOAL.copy()
        self.door_qadr = int(model.jnt_qposadr[model.names["joint"].index("doorjoint")])

    def reset_hand(self, steps=50):
        for _ in range(steps):
            self.e.mocap_pos[:] = self.HAND_INIT
            self.e.mocap_quat[:] = [1, 0, 1, 0]
            self.e.ctrl[:] = [-1, 1]
            self.e.step(self.FRAME_SKIP)

    def reset(self, door_angle=-np.pi / 3):
        self.e.reset()
        self.reset_hand()
        self.e.qpos[self.door_qadr] = door_angle
        self.e.qvel[self.door_qadr] = 0
        self.e.forward()
        return self.obs()

    def obs(self):
        e = self.e
        hand = e.site_xpos("body:hand")
        grip = np.clip(np.linalg.norm(e.site_xpos("rightEndEffector") - e.site_xpos("leftEndEffector")) / 0.1, 0.0, 1.0)
        return np.concatenate([hand, [grip], e.geom_xpos("handle"), self.goal])

    @staticmethod
    def _tolerance_gaussian(x, upper, margin, value_at_margin=0.1):
        """metaworld reward_utils.tolerance(x, bounds=(0, upper), margin, sigmoid='gaussian') [metaworld@master, not under
        /root/reference: formula as recalled in SURVEY.md Appendix C -- parity of the dense reward is UNPINNED]."""
        if 0 <= x <= upper:
            return 1.0
        if margin == 0:
            return 0.0
        d = (x - upper if x > upper else -x) / margin
        scale = np.sqrt(-2 * np.log(value_at_margin))
        return float(np.exp(-0.5 * (d * scale) ** 2))

    def dense_reward(self, obs):
        """compute_reward(obs)[0] with reward_type='dense' (reference earl_benchmark/envs/sawyer_door.py:141-171)."""
        tcp, obj, target = obs[:3], obs[4:7], obs[11:14]
        obj_init_pos = np.array([0.1, 0.95, 0.1], dtype=np.float32)
        tcp_to_obj, obj_to_target = np.linalg.norm(tcp - obj), np.linalg.norm(obj - target)
        in_place = self._tolerance_gaussian(obj_to_target, 0.05, np.linalg.norm(obj_init_pos - target))
        hand_in_place = self._tolerance_gaussian(tcp_to_obj, 0.25 * 0.05, np.linalg.norm(self.HAND_INIT - obj) + 0.1)
        reward = 3 * hand_in_place + 6 * in_place
        return 10.0 if obj_to_target < 0.05 else reward

    def step(self, action):
        a = np.clip(np.asarray(action, np.float64), -1, 1)
        self.e.mocap_pos[:] = np.clip(self.e.mocap_pos + a[:3] * self.ACTION_SCALE, self.MOCAP_LOW, self.MOCAP_HIGH)
        self.e.mocap_quat[:] = [1, 0, 1, 0]
        self.e.ctrl[:] = [a[3], -a[3]]
        self.e.step(self.FRAME_SKIP)
        # NO forward() here: mj_step integrates AFTER its mj_forward, so the body / site / geom poses the env reads right
        # after sim.step() are those of the state BEFORE the last substep's integration (one substep stale).
        o = self.obs()
        reward = float(np.linalg.norm(o[4:7] - o[11:14]) <= 0.02)
        return o, reward


class SawyerPegOracle:
    """metaworld SawyerXYZEnv.step / reset semantics + EARL SawyerPegV2 observation and sparse reward
    (reference earl_benchmark/envs/sawyer_peg.py:134-142,192-229,296-305)."""
    MOCAP_LOW = np.array([-0.5, 0.40, 0.05])
    MOCAP_HIGH = np.array([0.5, 1.0, 0.5])
    HAND_INIT = np.array([0, 0.6, 0.2], dtype=np.float64)
    GOAL = np.array([0.0, 0.6, 0.2, 1.0, -0.3 + 0.03, 0.6, 0.0 + 0.13])
    OBJ_INIT = np.array([0, 0.6, 0.02])
    FRAME_SKIP, ACTION_SCALE, TARGET_RADIUS = 5, 1.0 / 100, 0.05

    def __init__(self, model):
        self.e = Engine(model)
        self.goal = self.GOAL.copy()
        j = [k for k in range(len(model.jnt_type)) if model.jnt_type[k] == 0][0]
        self.peg_qadr, self.peg_dadr = int(model.jnt_qposadr[j]), int(model.jnt_dofadr[j])

    def reset_hand(self, steps=50):
        for _ in range(steps):
            self.e.mocap_pos[:] = self.HAND_INIT
            self.e.mocap_quat[:] = [1, 0, 1, 0]
            self.e.ctrl[:] = [-1, 1]
            self.e.step(self.FRAME_SKIP)

    def reset(self, peg_pos=None):
        self.e.reset()
        self.reset_hand()
        p = self.OBJ_INIT if peg_pos is None else np.asarray(peg_pos, np.float64)
        self.init_tcp = self._tcp_center()                        # _reset_hand: self.init_tcp = self.tcp_center
        self.e.qpos[self.peg_qadr:self.peg_qadr + 3] = p          # _set_obj_xyz: qpos[9:12] = pos, qvel[9:15] = 0
        self.e.qvel[self.peg_dadr:self.peg_dadr + 6] = 0
        self.e.forward()
        self.obj_init_pos = p.copy()                              # sawyer_peg.py:215-217
        self.peg_head_pos_init = self.e.site_xpos("pegHead")
        return self.obs()

    def obs(self):
        e = self.e
        hand = e.site_xpos("body:hand")
        grip = np.clip(np.linalg.norm(e.site_xpos("rightEndEffector") - e.site_xpos("leftEndEffector")) / 0.1, 0.0, 1.0)
        return np.concatenate([hand, [grip], e.site_xpos("pegHead"), self.goal])

    # ---- dense reward (reference earl_benchmark/envs/sawyer_peg.py:231-299).  metaworld's reward_utils and
    # SawyerXYZEnv._gripper_caging_reward are NOT under /root/reference: restated from SURVEY.md Appendix C and the
    # published metaworld sources (parity unpinned).
    @staticmethod
    def _tolerance_long_tail(x, lower, upper, margin, value_at_margin=0.1):
        if lower <= x <= upper:
            return 1.0
        if margin == 0:
            return 0.0
        d = (lower - x if x < lower else x - upper) / margin
        scale = np.sqrt(1 / value_at_margin - 1)
        return 1 / ((d * scale) ** 2 + 1)

    @staticmethod
    def _hamacher(a, b):
        den = a + b - a * b
        return a * b / den if den > 0 else 0.0

    @staticmethod
    def _rect_prism_tolerance(curr, zero, one):
        def in_range(a, b, c):
            return b <= a <= c if c >= b else c <= a <= b
        if all(in_range(curr[k], zero[k], one[k]) for k in range(3)):
            diff = one - zero
            return float(np.prod((curr - zero) / diff))
        return 1.0

    def _tcp_center(self):
        return 0.5 * (self.e.site_xpos("rightEndEffector") + self.e.site_xpos("leftEndEffector"))

    def _caging(self, action, obj_pos, obj_radius=0.0075, pad_success_thresh=0.03, xz_thresh=0.005):
        e = self.e
        pad_y = np.array([e.site_xpos("body:leftpad")[1], e.site_xpos("body:rightpad")[1]])
        to_obj, to_init = np.abs(pad_y - obj_pos[1]), np.abs(pad_y - self.obj_init_pos[1])
        margin = np.abs(to_init - pad_success_thresh)
        cy = [self._tolerance_long_tail(to_obj[i], obj_radius, pad_success_thresh, margin[i]) for i in range(2)]
        caging_y = self._hamacher(*cy)
        tcp = self._tcp_center()
        xz_margin = max(np.linalg.norm(self.obj_init_pos[[0, 2]] - self.init_tcp[[0, 2]]) - xz_thresh, 0.0)
        caging_xz = self._tolerance_long_tail(np.linalg.norm(tcp[[0, 2]] - obj_pos[[0, 2]]), 0.0, xz_thresh, xz_margin)
        closed = min(max(0.0, float(action[-1])), 1.0)
        caging = self._hamacher(caging_y, caging_xz)
        gripping = closed if caging > 0.97 else 0.0
        return 0.5 * (self._hamacher(caging, gripping) + caging)        # high_density=True

    def dense_reward(self, obs, action):
        e = self.e
        tcp, tcp_opened, obj_head, target = obs[:3], obs[3], obs[4:7], obs[11:14]
        obj = obs[4:7] - e.site_xpos("pegHead") + e.site_xpos("pegGrasp")
        tcp_to_obj = np.linalg.norm(obj - tcp)
        scale = np.array([1.0, 2.0, 2.0])
        obj_to_target = np.linalg.norm((obj_head - target) * scale)
        in_place = self._tolerance_long_tail(obj_to_target, 0.0, self.TARGET_RADIUS, np.linalg.norm((self.peg_head_pos_init - target) * scale))
        cb = [self._rect_prism_tolerance(obj_head, e.site_xpos(f"bottom_right_corner_collision_box_{q}"),
                                         e.site_xpos(f"top_left_corner_collision_box_{q}")) for q in (1, 2)]
        in_place = self._hamacher(in_place, self._hamacher(cb[1], cb[0]))
        lifted = tcp_to_obj < 0.08 and tcp_opened > 0 and obj[2] - 0.01 > self.obj_init_pos[2]
        grasped = 1.0 if lifted else self._caging(action, obj)
        reward = self._hamacher(grasped, in_place)
        if lifted:
            reward += 1.0 + 5 * in_place
        if obj_to_target <= self.TARGET_RADIUS:
            reward = 10.0
        return reward

    def step(self, action):
        a = np.clip(np.asarray(action, np.float64), -1, 1)
        self.e.mocap_pos[:] = np.clip(self.e.mocap_pos + a[:3] * self.ACTION_SCALE, self.MOCAP_LOW, self.MOCAP_HIGH)
        self.e.mocap_quat[:] = [1, 0, 1, 0]
        self.e.ctrl[:] = [a[3], -a[3]]
        self.e.step(self.FRAME_SKIP)
        o = self.obs()  # stale by one substep, as in the reference (see SawyerDoorOracle.step)
        return o, float(np.linalg.norm(o[4:7] - o[11:14]) <= self.TARGET_RADIUS)


class KitchenOracle:
    """Reference kitchen task on the fp64 checker engine: ENV/kitchen.py::Kitchen over adept_envs KitchenV0 / Robot_VelAct
    (SURVEY.md 3.4, rows a12-a14).  The logic around the physics is oracle/kitchen_logic.py (pinned bit for bit against
    the reference's own code); the physics is this file's Engine on the compiled franka_kitchen_jntpos_act_ab.xml model
    (friction-loss rows, joint equalities, pyramidal cones, capsule collisions) -- PARITY UNPINNED: the reference ships no
    kitchen trajectory, MuJoCo is not in this image."""

    def __init__(self, model):
        from . import kitchen_logic as KL
        self.KL = KL
        self.e = Engine(model)
        self.logic = KL.KitchenLogic()
        self.site_names = KL.SITES

    def seed(self, seed):
        self.logic.seed(seed)

    def _simulate(self, ctrl):
        self.e.ctrl[:] = ctrl                       # do_simulation writes ctrl[0:nu]; the engine clamps to ctrlrange
        self.e.step(self.KL.FRAME_SKIP)

    def sites(self):
        return np.stack([self.e.site_xpos(s) for s in self.site_names])

    def reset(self):
        """Kitchen.reset_model (ENV/kitchen.py:118-139): drawn object configuration, robot.reset (sim.reset + qpos write +
        forward + 5 cached observations at noise ratio 1), mocap <- midpoint, 10 x robot.step(0), observation."""
        KL = self.KL
        q0, self.config_index = self.logic.reset_state()
        self.e.reset()
        self.e.qpos[:] = q0
        self.e.qvel[:] = 0
        self.e.forward()
        for _ in range(5):
            self.logic.observe(self.e.qpos, noise_ratio=1)
        self.e.mocap_pos[:] = KL.MIDPOINT
        for _ in range(10):
            _, ctrl = self.logic.control(np.zeros(9), KL.MIDPOINT)
            self._simulate(ctrl)
        return self.logic.observe(self.e.qpos)

    def step(self, action):
        """KitchenV0.step (ADEPT/franka/kitchen_multitask_v0.py:91-125) -> (obs, reward, success)."""
        mocap, ctrl = self.logic.control(np.asarray(action, np.float64), self.e.mocap_pos.copy())
        self.e.mocap_pos[:] = mocap
        self._simulate(ctrl)
        obs = self.logic.observe(self.e.qpos)
        return obs, self.logic.reward(obs, self.e.mocap_pos.copy(), self.sites()), self.logic.success(obs)
