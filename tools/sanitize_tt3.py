"""compute-sanitizer target for the three-object tabletop: ragged batch, device + host path, masked resets, rollout ring,
reward / state entry points.  Usage: compute-sanitizer --tool memcheck python tools/sanitize_tt3.py"""
import numpy as np
import torch

from earl_benchmark_b200.envs.tabletop_manipulation_3obj import TabletopManipulation
from earl_benchmark_b200.wrappers.persistent_state_wrapper import PersistentStateWrapper

for n, reward_type in ((1, "sparse"), (257, "dense"), (300003, "sparse")):
    env = PersistentStateWrapper(TabletopManipulation(reward_type=reward_type, num_envs=n, device="cuda:0"), 7)
    rs = np.random.RandomState(n)
    q0 = np.concatenate([rs.uniform(-2, 2, (n, 2))] * 4, axis=1)
    q0[:, 2:] += rs.uniform(-0.5, 0.5, (n, 6))
    env.reset(init_qpos=q0)
    for t in range(10):
        a = rs.uniform(-1.2, 1.2, (n, 3)).astype(np.float32)
        a[: n // 2, 2] = 1.0
        o, r, d, info = env.step(torch.from_numpy(a).cuda()) if t % 2 == 0 else env.step(a)
        d = d.cpu().numpy() if isinstance(d, torch.Tensor) else d
        if d.any():
            env.reset(mask=d)
    acts = torch.rand((3, n, 3), device="cuda") * 2 - 1
    obs = torch.empty((2, n, 20), device="cuda")
    rew = torch.empty((2, n), device="cuda")
    done = torch.empty((2, n), dtype=torch.uint8, device="cuda")
    env.rollout_into(acts, 7, obs, rew, done)     # ring slots of a ragged batch are only 4-byte aligned: scalar action loads
    env.compute_reward(env._get_obs())
    env.is_successful()
    q, att = env.get_state()
    env.set_state(qpos=q, attached=att.cpu().numpy())
    env.reset_goal()
    torch.cuda.synchronize()
    print(n, reward_type, "ok", env.total_steps, int(env.num_interventions.max()), env.launch_count)
