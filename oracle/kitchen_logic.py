"""CPU restatement of the reference's kitchen task logic AROUND the physics -- TEST INFRASTRUCTURE ONLY.

Covers SURVEY.md section 8 rows a12 (the glue of KitchenV0.step: everything except the 40 x mj_step), a13 (noisy
observation) and a14 (reward / success), plus the reset draw.  Pinned bit for bit against the reference's own code run
behind a scripted MuJoCo stand-in (oracle/gen_kitchen_golden.py -> tests/golden/kitchen_ref_logic.npz,
tests/test_kitchen_logic.py).  Paths: ENV/ = earl_benchmark/envs/, ADEPT/ = ENV/kitchen_assets/adept_envs/adept_envs/.
"""
import numpy as np

N_ROBOT, N_OBJ = 9, 14                      # ADEPT/franka/kitchen_multitask_v0.py:33-34
# ENV/kitchen.py:15-25
COMPONENTS = (("burner0", (9, 10)), ("burner1", (11, 12)), ("burner2", (13, 14)), ("burner3", (15, 16)),
              ("light_switch", (17, 18)), ("slide_cabinet", (19,)), ("hinge_cabinet", (20, 21)), ("microwave", (22,)))
# ENV/kitchen.py:149-156
TASK_SITE = {"microwave": "microhandle_site", "hinge_cabinet": "hinge_site2", "slide_cabinet": "slide_site",
             "burner0": "knob1_site", "burner1": "knob2_site", "burner2": "knob3_site", "burner3": "knob4_site",
             "light_switch": "light_site"}
SITES = ("microhandle_site", "hinge_site2", "slide_site", "knob1_site", "knob2_site", "knob3_site", "knob4_site", "light_site")
# ENV/kitchen.py:28-52
GOAL = np.array([-4.1336253e-01, -1.6970085e+00, 1.4286385e+00, -2.5005307e+00, 6.2198675e-01, 1.2632011e+00, 8.8903642e-01,
                 4.3514766e-02, 7.9217982e-03, -5.1586074e-04, 4.8548312e-04, -5.4527864e-06, 6.3510129e-06, 6.0837720e-05,
                 -3.3861103e-05, 6.6394619e-05, -1.9801613e-05, -1.2477605e-04, 3.8065159e-04, -1.5148541e-04, -9.2229841e-04,
                 7.2293887e-03, 6.9650509e-03])
# ADEPT/franka/kitchen_multitask_v0.py:65-70
INIT_QPOS = np.array([1.48388023e-01, -1.76848573e+00, 1.84390296e+00, -2.47685760e+00, 2.60252026e-01, 7.12533105e-01,
                      1.59515394e+00, 4.79267505e-02, 3.71350919e-02, -2.66279850e-04, -5.18043486e-05, 3.12877220e-05,
                      -4.51199853e-05, -3.90842156e-06, -4.22629655e-05, 6.28065475e-05, 4.04984708e-05, 4.62730939e-04,
                      -2.26906415e-04, -4.65501369e-04, -6.44129196e-03, -1.77048263e-03, 1.08009684e-03])
MIDPOINT = np.array([-0.440, 0.1, 2.226])                  # :44
MOCAP_LOW, MOCAP_HIGH = np.array([-0.7, -0.1, 1.8]), np.array([0.4, 0.5, 2.6])   # :47-48
MOCAP_RANGE = np.array([0.01, 0.01, 0.01])                 # :45
FRAME_SKIP, TIMESTEP = 40, 0.002
NOISE_RATIO = 0.1                                          # :41
# ADEPT/franka/robot/franka_config.xml:17-45 (qpos0..qpos22): pos_bound, vel_bound, pos_noise_amp
POS_BOUND = np.array([[-2.9, 2.9], [-1.8, 1.8], [-2.9, 2.9], [-3.1, 0.0], [-2.9, 2.9], [0.0, 3.8], [-2.9, 2.9], [0.0, 0.04], [0.0, 0.04]]
                     + [[-.5, 0.0]] * 2 + [[-.005, 0.0]] * 6 + [[-1.5, 1.5]] * 3 + [[-10.57, 10.57]] * 3)
VEL_BOUND = np.array([[-10.0, 10.0]] * 9 + [[-5.0, 5.0]] * 11 + [[-.5, .5]] * 3)
POS_NOISE_AMP = np.array([0.1] * 9 + [0.005] * 2 + [0.0005] * 6 + [0.005] * 3 + [0.1] * 3)
VEL_NOISE_AMP = np.array([0.1] * 9 + [0.005] * 2 + [0.005] * 6 + [0.005] * 3 + [0.1] * 3)


def _initial_states():
    """ENV/kitchen.py:59-85: the six two-object configurations ('all_pairs')."""
    val = {"microwave": ((22,), [-0.7]), "light_switch": ((17, 18), [-0.69, -0.05]), "slide_cabinet": ((19,), [0.37]),
           "hinge_cabinet": ((20, 21), [0., 1.45])}
    rows = []
    for pair in (("microwave", "hinge_cabinet"), ("microwave", "slide_cabinet"), ("microwave", "light_switch"),
                 ("light_switch", "slide_cabinet"), ("light_switch", "hinge_cabinet"), ("slide_cabinet", "hinge_cabinet")):
        s = GOAL.copy()
        for name in pair:
            s[list(val[name][0])] = np.array(val[name][1])
        rows.append(s)
    return np.array(rows)


ALL_PAIRS = _initial_states()


class KitchenLogic:
    """Everything KitchenV0.step / reset do except advancing the physics."""

    def __init__(self):
        self.goal = GOAL.copy()
        self.np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(None)))
        self.last_qp = np.zeros(N_ROBOT)

    def seed(self, seed):
        """MujocoEnv._seed (ADEPT/mujoco_env.py:113-118) -> gym 0.23.1 seeding.np_random: PCG64(SeedSequence(seed))."""
        self.np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))

    def observe(self, qpos, noise_ratio=NOISE_RATIO):
        """Robot.get_obs (ADEPT/franka/robot/franka_robot.py:137-168) + KitchenV0._get_obs (:127-139).  Four draws per
        observation, in this order: robot qpos (9), robot qvel (9), object qpos (14), object qvel (14); the velocity
        draws only advance the stream.  KitchenV0._get_obs passes robot_noise_ratio = 0.1; the cache refresh of
        Robot.reset (_observation_cache_refresh, :120-122) calls get_obs with its DEFAULT ratio 1, so the controls of
        the ten settle steps of a reset come from an observation with ten times the noise."""
        u = self.np_random.uniform
        qp = np.array(qpos[:N_ROBOT], dtype=np.float64)
        qp += noise_ratio * POS_NOISE_AMP[:N_ROBOT] * u(low=-1., high=1., size=N_ROBOT)
        u(low=-1., high=1., size=N_ROBOT)
        obj = np.array(qpos[-N_OBJ:], dtype=np.float64)
        obj += noise_ratio * POS_NOISE_AMP[-N_OBJ:] * u(low=-1., high=1., size=N_OBJ)
        u(low=-1., high=1., size=N_OBJ)
        self.last_qp = qp
        return np.concatenate([qp, obj, self.goal])

    def control(self, action, mocap_pos):
        """KitchenV0.step :91-105 and Robot.step / Robot_VelAct.ctrl_velocity_limits / ctrl_position_limits
        (franka_robot.py:172-207,255-264): returns (new mocap position, the nu = 2 controls the simulation sees)."""
        a = np.clip(action, -1.0, 1.0)
        a = 0.0 + a * 2.0                                           # act_mid + a * act_amp
        new_mocap = np.clip(mocap_pos + a[:3] * MOCAP_RANGE, MOCAP_LOW, MOCAP_HIGH)
        vel = np.clip(a, VEL_BOUND[:N_ROBOT, 0], VEL_BOUND[:N_ROBOT, 1])
        pos = self.last_qp + vel * (FRAME_SKIP * TIMESTEP)
        pos = np.clip(pos, POS_BOUND[:N_ROBOT, 0], POS_BOUND[:N_ROBOT, 1])
        return new_mocap, pos[:2]                                  # do_simulation writes ctrl[0:nu] (ADEPT/mujoco_env.py:148-153)

    def reward(self, obs, mocap_pos, site_xpos):
        """Kitchen._get_reward_n_score (ENV/kitchen.py:141-175); site_xpos: dict name -> xyz or [8,3] in SITES order."""
        if not isinstance(site_xpos, dict):
            site_xpos = dict(zip(SITES, site_xpos))
        r = -10 * np.linalg.norm(obs[9:23] - obs[9 + 23:23 + 23])
        reaching = False
        for key, idx in COMPONENTS:
            cur = np.array(idx)
            if np.linalg.norm(obs[cur] - obs[cur + 23]) < len(idx) * 0.01:
                r += 1
            elif not reaching:
                reaching = True
                r += -0.5 * np.linalg.norm(mocap_pos - site_xpos[TASK_SITE[key]])
        return r

    @staticmethod
    def success(obs):
        """Kitchen.is_successful (ENV/kitchen.py:181-183)."""
        return bool(np.linalg.norm(obs[9:23] - obs[9 + 23:23 + 23]) <= 0.3)

    @staticmethod
    def reset_state():
        """Kitchen.reset_model (ENV/kitchen.py:118-127): one draw from the GLOBAL legacy np.random stream."""
        q = INIT_QPOS.copy()
        idx = np.random.randint(ALL_PAIRS.shape[0])
        q[9:] = ALL_PAIRS[idx, 9:]
        return q, idx
