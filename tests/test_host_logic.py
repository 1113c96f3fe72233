"""Host-side mirror of the reference loader surface (earl_benchmark/__init__.py) -- no GPU needed."""
import os

import numpy as np
import pytest

import earl_benchmark_b200 as eb
from earl_benchmark_b200 import demos, shard_range
from earl_benchmark_b200.wrappers import lifelong_wrapper, persistent_state_wrapper


def test_config_tables_and_horizons():
    e = eb.EARLEnvs("tabletop_manipulation", reward_type="sparse", num_envs=8)
    tr, ev = e.get_envs()
    assert isinstance(tr, persistent_state_wrapper.PersistentStateWrapper)
    assert tr._episode_horizon == 200000 and ev._episode_horizon == 200
    assert tr.env is not ev.env  # two independent env instances (reference :104-105)
    e = eb.EARLEnvs("tabletop_manipulation", train_horizon=123, eval_horizon=7)
    tr, ev = e.get_envs()
    assert tr.env._episode_horizon == 123 and ev.env._episode_horizon == 7
    e = eb.EARLEnvs("tabletop_manipulation", setup_as_lifelong_learning=True)
    ll = e.get_envs()
    assert isinstance(ll, lifelong_wrapper.LifelongWrapper)
    assert ll._goal_change_frequency == 400 and ll.env._episode_horizon == 50000
    assert ll._base._lifelong and ll._base._goal_change_frequency == 400


def test_unknown_env_is_keyerror_like_reference():
    with pytest.raises(KeyError):
        eb.EARLEnvs("no_such_env")


def test_every_task_of_the_loader_is_built():
    """All four tasks of EARLEnvs construct (lazily: no device work before the first reset), dense peg reward included."""
    for name, rt in (("tabletop_manipulation", "sparse"), ("sawyer_door", "dense"), ("sawyer_peg", "dense"), ("kitchen", "dense")):
        tr, ev = eb.EARLEnvs(name, reward_type=rt, num_envs=2).get_envs()
        assert tr.num_envs == 2


def test_kitchen_loader_surface():
    """Reference: Kitchen(reward_type='sparse') raises ValueError (envs/kitchen.py:91-92) and EARLEnvs passes its default
    'sparse' through (__init__.py:135-138), so the dense reward has to be asked for; initial / goal states :205-236."""
    with pytest.raises(ValueError):
        eb.EARLEnvs("kitchen")
    e = eb.EARLEnvs("kitchen", reward_type="dense", num_envs=3)
    assert e.get_initial_states().shape == (6, 23) and e.get_goal_states().shape == (1, 23)
    tr, ev = e.get_envs()
    assert tr.get_task() == "all_pairs" and tr.get_init_states().shape == (6, 23)
    assert np.array_equal(tr.get_next_goal(), e.get_goal_states()[0])
    assert not e.has_demos()
    from earl_benchmark_b200.envs import kitchen
    st = kitchen.pcg64_states([5, 6])
    g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(5)))
    s = g.bit_generator.state["state"]
    assert int(st[0, 0]) == s["state"] >> 64 and int(st[0, 1]) == s["state"] & ((1 << 64) - 1) and int(st[0, 3]) == s["inc"] & ((1 << 64) - 1)
    import ctypes
    assert ctypes.sizeof(kitchen.MjkConfig) == 984


def test_states_and_demos(golden_dir):
    e = eb.EARLEnvs("tabletop_manipulation")
    g = np.load(os.path.join(golden_dir, "loader_constants.npz"))
    assert np.array_equal(e.get_initial_states(), g["tabletop_manipulation_initial_states"])
    assert np.array_equal(e.get_goal_states(), g["tabletop_manipulation_goal_states"])
    assert e.has_demos()
    fwd, rev = e.get_demonstrations()
    assert set(fwd) == {"observations", "actions", "rewards", "terminals", "next_observations", "infos"}
    assert fwd["observations"].shape == (1278, 12) and rev["actions"].shape == (1256, 3)
    assert fwd["terminals"].sum() == 12 and fwd["rewards"].dtype == np.float32
    for env, (nf, nr, od, ad) in {"sawyer_door": (395, 700, 14, 4), "sawyer_peg": (683, 1132, 14, 4)}.items():
        assert demos.load(env, "forward")["observations"].shape == (nf, od)
        assert demos.load(env, "reverse")["actions"].shape == (nr, ad)


def test_shard_range_partitions_exactly():
    for n in (1, 7, 8, 65536, 1000003):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_goal_table_and_custom_goals():
    from earl_benchmark_b200.envs.tabletop_manipulation import TabletopManipulation, goal_states, initial_states
    env = TabletopManipulation(reward_type="sparse", num_envs=3)
    assert len(env._goal_table) == 5 and np.array_equal(env._goal_table[4], initial_states[0])
    assert np.array_equal(env._goal_table[2][2:4], goal_states[2][2:4]) and np.array_equal(env._goal_table[2][:2], [0, 0])
    assert env._task_to_row.tolist() == [0, 3, 1, 2]
    rows = env._goal_rows_for(np.array([0.0, 0.0, 2.5, 0.0, -1, -1]))
    assert rows.tolist() == [4, 4, 4]
    rows = env._goal_rows_for(np.array([[0, 0, -2.5, 1, -1, -1], [0.5, 0, 1, 1, -1, -1], [0, 0, 2.5, 0, -1, -1]], float))
    assert rows.tolist() == [1, 5, 4] and len(env._goal_table) == 6
    with pytest.raises(ValueError):
        TabletopManipulation(task_list="bc_r")
    with pytest.raises(ValueError):
        TabletopManipulation(reward_type="banana")


def test_peg_reset_draws_replicate_the_reference_numpy_stream():
    """reset_model() of sawyer_peg.py:192-229 against numpy's own legacy generator, for the three variants."""
    from earl_benchmark_b200.envs import sawyer_peg as sp
    low, high = sp._RESET_LOW, sp._RESET_HIGH
    pos_box = sp.goal_states[0][4:] - np.array([0.03, 0.0, 0.13])

    def ref_draws(seed, count, reset_at_goal, wide_init):
        rs = np.random.RandomState(seed)
        out = []
        for _ in range(count):
            if not reset_at_goal:
                rs.randint(0, 1)
                row, pos = 0, np.array([0, 0.6, 0.02])
                sample = lambda: np.split(rs.uniform(low, high, size=6), 2)[0]  # noqa: E731
                if wide_init:
                    if rs.uniform() < 0.5:
                        pos = sample()
                        while np.linalg.norm(pos[:2] - pos_box[:2]) < 0.1:
                            pos = sample()
                    else:
                        pos = sp.wide_initial_states[rs.randint(0, sp.wide_initial_states.shape[0])] - np.array([-0.1, 0., 0.])
                        pos = pos + rs.uniform(-0.02, 0.02, size=3)
                else:
                    pos = sample()
                    while np.linalg.norm(pos[:2] - pos_box[:2]) < 0.1:
                        pos = sample()
            else:
                row = 1 + rs.randint(0, sp.initial_states.shape[0])
                pos = sp.goal_states[0][4:] - np.array([-0.1, 0., 0.]) + rs.uniform(-0.02, 0.02, size=3)
            out.append((row, pos))
        return out

    for kw in (dict(), dict(wide_init=True), dict(reset_at_goal=True)):
        env = sp.SawyerPegV2(reward_type="sparse", num_envs=1, seed=7, **kw)
        want = ref_draws(7, 40, kw.get("reset_at_goal", False), kw.get("wide_init", False))
        for row, pos in want:
            r, p = env._draw_one()
            assert r == row and np.array_equal(p, pos)


def test_demo_utilities():
    fwd = demos.load("sawyer_door", "forward")
    eps = demos.episodes(fwd)
    assert [b - a for a, b in eps] == [78, 79, 75, 78, 85]
    ang = demos.door_angle_from_obs(fwd["observations"][[a for a, _ in eps]])
    assert np.all(ang > -np.pi / 3 - 1e-3) and np.all(ang < -np.pi / 3 + np.pi / 20 + 1e-3)   # reset_model's draw range
    dev = demos.load_to_device("sawyer_door", "forward", "cpu")
    assert dev["observations"].shape == (395, 14) and str(dev["actions"].dtype) == "torch.float32"
    peg = demos.load("sawyer_peg", "forward")
    p = demos.peg_position_from_obs(peg["observations"][0])
    assert 0.0 <= p[0] <= 0.2 and 0.5 <= p[1] <= 0.7 and abs(p[2] - 0.02) < 1e-6                # obj_low / obj_high


def test_sawyer_reset_draws_do_not_depend_on_the_number_of_shards():
    """A job of 10 envs sharded over 1, 2 or 3 GPUs draws the same door angles / peg positions for the same global env
    (every shard consumes the global np.random stream and keeps its slice)."""
    from earl_benchmark_b200.envs import sawyer_door as sd, sawyer_peg as sp
    total = 10
    whole = sd.SawyerDoorV2(num_envs=total, seed=5)._draw_angles(None)
    assert np.array_equal(whole, -np.pi / 3 + np.random.RandomState(5).uniform(0, np.pi / 20, total))
    for world in (2, 3):
        parts = []
        for r in range(world):
            lo, hi = shard_range(total, r, world)
            parts.append(sd.SawyerDoorV2(num_envs=hi - lo, seed=5, env_offset=lo, total_envs=total)._draw_angles(None))
        assert np.array_equal(np.concatenate(parts), whole)

    def peg_draws(n, off):
        env = sp.SawyerPegV2(reward_type="sparse", num_envs=n, seed=5, env_offset=off, total_envs=total)
        out = []
        for g in range(total):            # the loop of SawyerPegV2.reset(mask=None)
            row, p = env._draw_one()
            if off <= g < off + n:
                out.append(p)
        return np.array(out)
    whole_p = peg_draws(total, 0)
    lo, hi = shard_range(total, 1, 2)
    assert np.array_equal(peg_draws(hi - lo, lo), whole_p[lo:hi])
    # kitchen: np.random.randint(6) per env and one PCG64 observation-noise stream per GLOBAL env index
    from earl_benchmark_b200.envs import kitchen as kt
    whole_k = kt.Kitchen(num_envs=total, seed=5)._draw_configs(total)[1]
    np.random.seed(5)
    assert whole_k.tolist() == [np.random.randint(6) for _ in range(total)]
    for world in (2, 3):
        parts = []
        for r in range(world):
            lo, hi = shard_range(total, r, world)
            parts.append(kt.Kitchen(num_envs=hi - lo, seed=5, env_offset=lo, total_envs=total)._draw_configs(hi - lo)[1])
        assert np.array_equal(np.concatenate(parts), whole_k)


def test_three_object_env_host_side_contract():
    """Constructor validation, goal-table bookkeeping and wrapper configuration of the three-object env need no device."""
    import pytest
    from earl_benchmark_b200.envs import tabletop_manipulation_3obj as t3
    from earl_benchmark_b200.wrappers.lifelong_wrapper import LifelongWrapper
    from earl_benchmark_b200.wrappers.persistent_state_wrapper import PersistentStateWrapper
    with pytest.raises(ValueError):
        t3.TabletopManipulation(reward_type="shaped")
    with pytest.raises(ValueError):
        t3.TabletopManipulation(device="cpu")
    env = t3.TabletopManipulation(reward_type="sparse", num_envs=3)
    assert env.action_space.shape == (3,) and env.observation_space.shape == (20,)
    assert env.object_dict == {(0, 0): [2, 3], (0.5, 0.5): [4, 5], (1, 1): [6, 7]}     # reference :31-35
    assert np.array_equal(env.get_next_goal(), np.broadcast_to(t3.goal_states[0], (3, 10)))
    assert list(env._rows_for(t3.goal_states[0])) == [0, 0, 0]
    custom = t3.goal_states[0].copy()
    custom[2:4] = [1.0, 1.0]
    assert list(env._rows_for(np.stack([t3.goal_states[0], custom, custom]))) == [0, 1, 1]   # appended once
    for k in range(14):
        c = custom.copy()
        c[4] = 0.1 * (k + 1)
        env._rows_for(c)
    with pytest.raises(ValueError):                                                     # 16 rows at most
        c = custom.copy()
        c[4] = -1.0
        env._rows_for(c)
    w = PersistentStateWrapper(t3.TabletopManipulation(num_envs=2), 123)
    assert w.env._episode_horizon == 123
    with pytest.raises(ValueError):
        LifelongWrapper(t3.TabletopManipulation(num_envs=2), 400)


def test_host_zerocopy_policy_follows_the_ranks_on_this_host(monkeypatch):
    """earl_set_host_zerocopy's Python side: alone on the host -> the step kernel drives the pinned buffers itself; several
    ranks on the node (torchrun) -> staged copy pipeline; EARL_TT_HOST_ZEROCOPY set -> the library's own reading decides."""
    from earl_benchmark_b200.envs import _hostio
    for k in ("LOCAL_WORLD_SIZE", "WORLD_SIZE", "EARL_TT_HOST_ZEROCOPY"):
        monkeypatch.delenv(k, raising=False)
    assert _hostio.ranks_on_this_host() == 1 and _hostio.host_zerocopy_default() == 1
    monkeypatch.setenv("WORLD_SIZE", "8")
    assert _hostio.ranks_on_this_host() == 8 and _hostio.host_zerocopy_default() == 0
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "1")                          # 8 nodes x 1 GPU: nobody shares this host
    assert _hostio.host_zerocopy_default() == 1
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "2")
    monkeypatch.setenv("EARL_TT_HOST_ZEROCOPY", "1")
    assert _hostio.host_zerocopy_default() is None
