"""Kitchen task on the device (envs/kitchen.py -> include/earl_mj_kitchen_b200.h) against the checker
(oracle/engine.py::KitchenOracle = pinned task logic + fp64 engine).  Tolerances: the observation NOISE and all integer
bookkeeping are exact; states go through 40 fp32 substeps per env step (north-star bar 1e-4 per one-step qpos)."""
import numpy as np
import pytest
import torch

import earl_benchmark_b200 as eb
from earl_benchmark_b200.envs import kitchen
from earl_benchmark_b200.mjcf.compile import Model
from oracle import kitchen_logic as KL
from oracle.engine import KitchenOracle

pytestmark = pytest.mark.gpu


def _oracle(seed):
    k = KitchenOracle(Model.load(kitchen.MODEL_PATH))
    k.seed(seed)
    return k


def test_reset_and_open_loop_rollout_against_checker():
    n, seed = 6, 11
    env = kitchen.Kitchen(num_envs=n, device="cuda:0", seed=seed)
    env.seed(seed)
    ob = env.reset(config_index=np.arange(n) % 6).cpu().numpy()
    oracles = []
    for i in range(n):
        k = _oracle(seed + i)
        q0 = KL.INIT_QPOS.copy()
        q0[9:] = KL.ALL_PAIRS[i % 6, 9:]
        # KitchenOracle.reset with a given configuration
        k.e.reset()
        k.e.qpos[:], k.e.qvel[:] = q0, 0
        k.e.forward()
        for _ in range(5):
            k.logic.observe(k.e.qpos, noise_ratio=1)
        k.e.mocap_pos[:] = KL.MIDPOINT
        for _ in range(10):
            _, ctrl = k.logic.control(np.zeros(9), KL.MIDPOINT)
            k._simulate(ctrl)
        ob_ref = k.logic.observe(k.e.qpos)
        oracles.append(k)
        assert np.abs(ob[i] - ob_ref).max() < 2e-5, (i, np.abs(ob[i] - ob_ref).max())
        assert np.array_equal(ob[i, 23:], KL.GOAL)
    st = env.get_state()
    for i, k in enumerate(oracles):
        assert np.array_equal(st["mocap_pos"][i], KL.MIDPOINT)
        assert np.abs(st["last_noisy_qp"][i] - k.logic.last_qp).max() < 2e-5
        # the noise itself is exact: obs - state on both sides is the same PCG64 draw times the same amplitude
        assert np.abs((ob[i, :9] - st["qpos"][i, :9]) - (k.logic.last_qp - k.e.qpos[:9])).max() < 2e-7
    rs = np.random.RandomState(3)
    worst_q = worst_ob = worst_r = 0.0
    for t in range(12):
        a = rs.uniform(-1.2, 1.2, (n, 9)).astype(np.float32)
        obs, rew, done, info = env.step(torch.from_numpy(a).cuda())
        obs, rew = obs.cpu().numpy(), rew.cpu().numpy()
        st = env.get_state()
        for i, k in enumerate(oracles):
            ob_ref, r_ref, s_ref = k.step(a[i])
            assert np.abs(st["mocap_pos"][i] - k.e.mocap_pos).max() == 0          # fp64 mocap arithmetic: exact
            worst_q = max(worst_q, np.abs(st["qpos"][i] - k.e.qpos).max())
            worst_ob = max(worst_ob, np.abs(obs[i] - ob_ref).max())
            worst_r = max(worst_r, abs(rew[i] - r_ref))
            assert bool(info["success"][i]) == s_ref and not bool(done[i])
    print("kitchen open loop, 12 env steps: |dqpos| %.2e |dobs| %.2e |dreward| %.2e" % (worst_q, worst_ob, worst_r))
    assert worst_q < 2e-3 and worst_ob < 2e-3 and worst_r < 2e-2
    assert env.work_counters()["bad_states"] == 0


def test_one_env_step_from_identical_states():
    """set_state on both sides (states along a scripted checker rollout into the cabinets), one env step = 40 substeps."""
    import test_kitchen_engine_gpu as T
    kref, states = T._states(24)
    n = len(states)
    env = kitchen.Kitchen(num_envs=n, device="cuda:0", seed=5)
    env.seed(5)
    env.reset(config_index=np.zeros(n, int))
    last = np.stack([s[0][:9] for s in states])
    env.set_state(qpos=np.stack([s[0] for s in states]), qvel=np.stack([s[1] for s in states]),
                  qacc_warmstart=np.stack([s[2] for s in states]), mocap_pos=np.stack([s[3] for s in states]), last_noisy_qp=last)
    rs = np.random.RandomState(9)
    a = rs.uniform(-1, 1, (n, 9)).astype(np.float32)
    obs, rew, done, info = env.step(torch.from_numpy(a).cuda())
    st = env.get_state()
    e = kref.e
    dq, dv = [], []
    for i, (q, v, w, mp, c) in enumerate(states):
        e.reset()
        e.qpos[:], e.qvel[:], e.mocap_pos[:] = q.astype(np.float32), v.astype(np.float32), mp
        e.arr("qacc_warmstart", (32,))[:23] = w.astype(np.float32)
        kref.logic.last_qp = last[i].copy()
        mocap, ctrl = kref.logic.control(a[i], e.mocap_pos.copy())
        e.mocap_pos[:] = mocap
        kref._simulate(ctrl)
        dq.append(np.abs(st["qpos"][i] - e.qpos).max())
        dv.append(np.abs(st["qvel"][i] - e.qvel).max())
        ref_sites = np.stack([e.site_xpos(nm) for nm in kitchen.REWARD_SITES])      # device order: component_to_state_idx
        # sites follow the joints (lever arms below 1 m): bounded by this state's own joint difference, which is judged below
        assert np.abs(st["site_xpos"][i] - ref_sites).max() < max(1e-5, 2 * dq[-1])
    dq, dv = np.array(dq), np.array(dv)
    print("kitchen one env step (40 substeps): dq median %.2e max %.2e, dv median %.2e max %.2e" % (np.median(dq), dq.max(), np.median(dv), dv.max()))
    # Two of the 24 states (finger capsules wedged against the slide cabinet) are UNSTABLE: the fp32 / fp64 difference doubles
    # every ~3 substeps (1e-7 -> 1e-4 in 20, measured per substep through the engine-level entry point, the host build of the
    # same source grows alike), then the contact sets part and the difference is macroscopic (1.3e-2 rad, 1.4 rad/s).  The
    # bulk is judged tightly, the tail only bounded.
    assert np.median(dq) < 1e-5 and np.percentile(dq, 90) < 2e-4 and dq.max() < 5e-2
    assert np.median(dv) < 1e-4 and np.percentile(dv, 90) < 2e-2 and dv.max() < 5.0


def test_loader_wrappers_counters_and_reset_draws():
    envs = eb.EARLEnvs("kitchen", reward_type="dense", num_envs=4, seed=21, train_horizon=3, eval_horizon=2)
    tr, ev = envs.get_envs()
    ob = tr.reset()
    assert tuple(ob.shape) == (4, 46) and ob.dtype == torch.float64
    # reset_model's np.random.randint(6), one legacy numpy stream in environment order
    np.random.seed(21)
    assert tr.last_config_index.tolist() == [np.random.randint(6) for _ in range(4)]
    assert tr.num_interventions.tolist() == [1] * 4 and tr.total_steps == 0
    a = torch.zeros((4, 9), device="cuda")
    for t in range(3):
        ob, r, done, info = tr.step(a)
        assert done.tolist() == [t == 2] * 4 and tr.steps_since_reset.tolist() == [t + 1] * 4
    assert tr.total_steps == 3
    ob_last, r_last = ob.clone(), r.clone()
    cr = tr.compute_reward(ob_last)
    assert np.abs(cr - r_last.cpu().numpy()).max() < 1e-9
    tr.reset(mask=[True, False, True, False])
    assert tr.num_interventions.tolist() == [2, 1, 2, 1] and tr.steps_since_reset.tolist() == [0, 3, 0, 3]
    assert bool(torch.all(tr.is_successful(ob) == info["success"]))
    ll = eb.EARLEnvs("kitchen", reward_type="dense", setup_as_lifelong_learning=True, num_envs=2, seed=2,
                     goal_change_frequency=2).get_envs()
    ll.seed(2)
    ll.reset(config_index=[0, 0])
    # LifelongWrapper on the checker: every 2nd step reset_goal() + env._get_obs() = a SECOND noisy observation, which is
    # returned and cached for the next control (lifelong_wrapper.py:36-42)
    k = _oracle(2)
    k.logic.reset_state = lambda: (np.concatenate([KL.INIT_QPOS[:9], KL.ALL_PAIRS[0, 9:]]), 0)
    k.reset()
    tot = torch.zeros(2, dtype=torch.float64, device="cuda")
    rs = np.random.RandomState(4)
    for t in range(5):
        a = rs.uniform(-1, 1, (2, 9)).astype(np.float32)
        ob, r, done, info = ll.step(torch.from_numpy(a).cuda())
        tot += r
        ob_ref, r_ref, _ = k.step(a[0])
        if t % 2 == 1:
            ob_ref = k.logic.observe(k.e.qpos)
        assert np.abs(ob[0].cpu().numpy() - ob_ref).max() < 1e-4 and abs(float(r[0]) - r_ref) < 1e-3
    assert torch.equal(ll.lifelong_return, tot)


def test_same_results_for_any_batch_composition():
    a = np.random.RandomState(1).uniform(-1, 1, (5, 9)).astype(np.float32)

    def run(idx):
        env = kitchen.Kitchen(num_envs=len(idx), device="cuda:0", seed=0)
        env.seed(0)
        # give every environment the stream of its GLOBAL index
        st = np.ascontiguousarray(kitchen.pcg64_states([int(i) for i in idx]))
        from earl_benchmark_b200 import _lib
        _lib.check(_lib.lib().earl_mjk_seed(env._handle, st.ctypes.data))
        env.reset(config_index=np.asarray(idx) % 6)
        out = []
        for t in range(3):
            ob, r, d, info = env.step(torch.from_numpy(np.repeat(a[t:t + 1], len(idx), 0)).cuda())
            out.append((ob.clone(), r.clone()))
        return out
    full, part = run([0, 1, 2, 3, 4, 5, 6]), run([5, 2])
    for (o1, r1), (o2, r2) in zip(full, part):
        assert torch.equal(o1[[5, 2]], o2) and torch.equal(r1[[5, 2]], r2)


def test_row_overflow_is_redone_by_the_extra_large_set_not_dropped(monkeypatch):
    """bench.py's kitchen stream: ~130 env steps in ~10^5 have a substep beyond the 112 rows of the primary set (a handful
    of six-dimensional finger contacts, 10 rows each).  With the redo pass they are re-stepped by the 544-row set
    (overflow_states stays 0); every environment that never overflowed is bit-identical to the run without it."""
    n, warm, steps = 14208, 30, 8   # bench.py's kitchen section: the arms have to get up to speed first

    def run(redo):
        monkeypatch.setenv("EARL_MJ_REDO", "1" if redo else "0")
        env = kitchen.Kitchen(num_envs=n, device="cuda:0", seed=0)
        env.seed(0)
        env.reset()
        gen = torch.Generator(device="cuda:0")
        gen.manual_seed(99)
        actions = torch.rand((8 + steps, n, 9), generator=gen, device="cuda:0", dtype=torch.float32) * 2 - 1
        rew = []
        for t in range(warm + steps):
            ob, r, d, info = env.step(actions[t % 8 if t < warm else 8 + t - warm])
            rew.append(r.clone())
        torch.cuda.synchronize()
        return env.get_state(), torch.stack(rew).cpu().numpy(), env.work_counters()

    st0, r0, w0 = run(False)
    st1, r1, w1 = run(True)
    assert w0["redone_states"] == 0 and w1["env_steps"] == w0["env_steps"] == n * (warm + steps)
    if w0["overflow_states"] == 0:
        pytest.skip("this action stream no longer outgrows the primary capacity set")
    assert w1["overflow_states"] == 0, w1
    assert 1 <= w1["redone_states"] <= 2 * w0["overflow_states"], (w0, w1)
    assert w1["bad_states"] == 0
    differ = np.flatnonzero((st0["qpos"] != st1["qpos"]).any(1) | (r0 != r1).any(0))
    assert 1 <= len(differ) <= w1["redone_states"], (len(differ), w1)
    assert np.isfinite(st1["qpos"]).all() and np.isfinite(r1).all()


def test_redo_scheduling_does_not_change_the_kitchen_physics(monkeypatch):
    """Contact-rich scripted policy (every arm driven into the slide-cabinet handle): a fifth of the env steps outgrow the
    112-row set.  Heavy-env routing, listing at the overflowing substep and the number of SMs left to the redo kernel are
    scheduling decisions: observations, rewards and states are bit-identical whether that number adapts or is fixed."""
    n, steps = 592, 70
    k = kitchen.REWARD_SITES.index("slide_site")

    def run(sms):
        if sms is None:
            monkeypatch.delenv("EARL_MJ_REDO_SMS", raising=False)
        else:
            monkeypatch.setenv("EARL_MJ_REDO_SMS", str(sms))
        env = kitchen.Kitchen(num_envs=n, device="cuda:0", seed=4)
        env.seed(4)
        env.reset()
        gen = torch.Generator(device="cuda:0")
        gen.manual_seed(8)
        rew = []
        for t in range(steps):
            st = env.get_state()
            d = torch.from_numpy(st["site_xpos"][:, k] - st["mocap_pos"]).to("cuda:0", torch.float32)
            u = torch.zeros((n, 9), device="cuda:0")
            u[:, :3] = torch.clamp(d * 10, -1, 1) * 0.5
            u += 0.1 * (torch.rand((n, 9), generator=gen, device="cuda:0") * 2 - 1)
            ob, r, dn, info = env.step(u)
            rew.append(r.clone())
        return ob.clone(), torch.stack(rew), env.get_state(), env.work_counters()

    o0, r0, s0, w0 = run(None)
    o1, r1, s1, w1 = run(6)
    if w0["redone_states"] == 0:
        pytest.skip("the scripted reach no longer outgrows the primary capacity set")
    assert w0["overflow_states"] == 0 and w1["overflow_states"] == 0 and w0["bad_states"] == 0
    assert w0["redone_states"] == w1["redone_states"] and w0["newton_iterations"] == w1["newton_iterations"]
    assert torch.equal(o0, o1) and torch.equal(r0, r1)
    assert np.array_equal(s0["qpos"], s1["qpos"]) and np.array_equal(s0["qvel"], s1["qvel"])
