"""Per-environment failure containment: metaworld catches MuJoCo's exception in do_simulation, stops
simulating that env and keeps returning the last stable observation with reward 0 until the next reset.  Here a step that
ends in a non-finite state is not stored: the env stays at its last good state, flag bit 2 is sticky until reset, the step
is counted in work_counters()['bad_states'] and every other env of the batch is untouched (missing #7)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_door_env_freezes_at_last_good_state_and_resumes_after_reset():
    from earl_benchmark_b200.envs import sawyer_door
    n, bad = 6, 2
    angles = -np.pi / 3 + np.linspace(0, np.pi / 20, n)
    rs = np.random.RandomState(3)
    acts = rs.uniform(-1, 1, (4, n, 4)).astype(np.float32)

    def run(inject):
        env = sawyer_door.SawyerDoorV2(num_envs=n, device="cuda:0")
        env.reset(door_angle=angles)
        env.step(torch.from_numpy(acts[0]).cuda())
        if inject:
            st = env.get_state()
            st["qvel"][bad, 3] = 3e38                         # finite in fp32, overflows inside the next substep
            env.set_state(qvel=st["qvel"])
        before = env._get_obs().cpu().numpy().copy()
        out = [[x.cpu().numpy().copy() for x in env.step(torch.from_numpy(a).cuda())[:3]] for a in acts[1:3]]
        return env, before, out

    ref_env, _, ref = run(False)
    env, before, out = run(True)
    for t in range(2):
        o, r, d = out[t]
        assert np.all(np.isfinite(o))
        assert np.array_equal(o[bad, :7], before[bad, :7])    # last stable observation, step after step
        assert r[bad] == 0.0
        keep = np.arange(n) != bad
        assert np.array_equal(o[keep], ref[t][0][keep]) and np.array_equal(r[keep], ref[t][1][keep])   # others untouched
    wc = env.work_counters()
    assert wc["bad_states"] == 2 and ref_env.work_counters()["bad_states"] == 0
    st = env.get_state()
    assert np.all(np.isfinite(st["qpos"])) and st["qvel"][bad, 3] > 1e38     # the injected (pre-step) state is what is kept
    # counters keep running like the wrapper's (steps are counted even though nothing moves); reset clears the flag
    mask = np.zeros(n, bool)
    mask[bad] = True
    env.reset(mask=mask, door_angle=angles)
    o, r, d, _ = env.step(torch.from_numpy(acts[3]).cuda())
    assert env.work_counters()["bad_states"] == 2
    fresh = sawyer_door.SawyerDoorV2(num_envs=n, device="cuda:0")
    fresh.reset(door_angle=angles)
    o2 = fresh.step(torch.from_numpy(acts[3]).cuda())[0]
    assert np.abs(o.cpu().numpy()[bad] - o2.cpu().numpy()[bad]).max() < 1e-6


def test_kitchen_failed_step_is_not_stored():
    from earl_benchmark_b200.envs import kitchen
    n, bad = 4, 1
    env = kitchen.Kitchen(num_envs=n, device="cuda:0", seed=2)
    env.seed(2)
    env.reset(config_index=np.arange(n))
    a = np.random.RandomState(0).uniform(-1, 1, (n, 9)).astype(np.float32)
    env.step(torch.from_numpy(a).cuda())
    st = env.get_state()
    good_qpos = st["qpos"].copy()
    st["qvel"][bad, 2] = 3e38
    env.set_state(qvel=st["qvel"])
    o, r, d, info = env.step(torch.from_numpy(a).cuda())
    o, r = o.cpu().numpy(), r.cpu().numpy()
    assert np.all(np.isfinite(o)) and r[bad] == 0.0 and np.all(r[np.arange(n) != bad] != 0.0)
    st2 = env.get_state()
    assert np.array_equal(st2["qpos"][bad], good_qpos[bad].astype(np.float32).astype(np.float64))    # not stored
    assert np.all(np.isfinite(st2["qpos"])) and env.work_counters()["bad_states"] == 1
    # the observation of the failed step is the (noisy) observation of the kept state: within the noise amplitude
    assert np.abs(o[bad, 9:23] - good_qpos[bad, 9:]).max() < 0.011
