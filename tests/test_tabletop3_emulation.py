"""The per-env arithmetic of the three-object tabletop KERNELS (earl_benchmark_b200/csrc/tt3_env.cuh, the header
csrc/earl_tt3.cu steps every env with) compiled for the host and checked, without a GPU, against outputs of the
UNMODIFIED reference class (tests/golden/tabletop3_ref_rollouts.npz) and against the numpy checker.  The GPU tests
(tests/test_tabletop3_gpu.py) check the same source on the device through the C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np

from conftest import REPO
from oracle import tabletop3

HERE = os.path.join(REPO, "tests", "host_emulation")
GOLD = np.load(os.path.join(REPO, "tests", "golden", "tabletop3_ref_rollouts.npz"))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so, src = os.path.join(HERE, "libemul_tt3.so"), os.path.join(HERE, "emul_tt3.cpp")
        hdr = os.path.join(REPO, "earl_benchmark_b200", "csrc", "tt3_env.cuh")
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in (src, hdr)):
            cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
            subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-ffp-contract=off", "-o", so, src])
        L = C.CDLL(so)
        DP = np.ctypeslib.ndpointer(np.float64, flags="C")
        FP = np.ctypeslib.ndpointer(np.float32, flags="C")
        L.emu_tt3_step.argtypes = [C.c_int, DP, np.ctypeslib.ndpointer(np.int32, flags="C"), FP, FP, C.c_int, C.c_double,
                                   C.c_double, C.c_double, C.c_double, FP, FP, np.ctypeslib.ndpointer(np.uint8, flags="C")]
        _LIB = L
    return _LIB


class Emu:
    """n envs stepped by the host build of the kernel arithmetic (reset and counters are not part of it)"""

    def __init__(self, qpos, dense=False):
        self.q = np.ascontiguousarray(qpos, np.float64).reshape(-1, 8).copy()
        self.n = len(self.q)
        self.att = np.zeros(self.n, np.int32)
        self.goal = np.ascontiguousarray(np.broadcast_to(tabletop3.GOAL_STATES[0].astype(np.float32), (self.n, 10)))
        self.dense = int(dense)

    def step(self, a):
        obs = np.empty((self.n, 20), np.float32)
        rew = np.empty(self.n, np.float32)
        suc = np.empty(self.n, np.uint8)
        lib().emu_tt3_step(self.n, self.q, self.att, np.ascontiguousarray(a, np.float32).reshape(self.n, 3), self.goal, self.dense,
                           0.4, 0.2, 2.8, 0.4, obs, rew, suc)
        return obs, rew, suc.astype(bool)


def gold(prefix):
    return {k[len(prefix) + 1:]: GOLD[k] for k in GOLD.files if k.startswith(prefix + "_")}


def replay(prefix, dense):
    g = gold(prefix)
    emu = Emu(g["qpos"][0], dense)
    for t in range(len(g["actions"])):
        o, r, s = emu.step(g["actions"][t])
        if g["reset_after"][t]:
            emu.q[0], emu.att[0] = g["qpos"][t + 1], 0             # the harness reset the reference env here
        else:
            assert np.array_equal(o[0], g["obs"][t + 1]), t
            assert np.array_equal(emu.q[0], g["qpos"][t + 1]) and emu.att[0] == g["attached"][t + 1], t
        if dense:
            assert abs(float(r[0]) - g["reward"][t]) <= 2e-6 * max(1.0, abs(g["reward"][t])), t
        else:
            assert r[0] == g["reward"][t], t
        assert bool(s[0]) == bool(g["success"][t]), t
    return g


def test_kernel_source_replays_the_reference_sparse_rollout():
    g = replay("sparse", dense=False)
    assert g["success"].sum() > 100 and set(np.unique(g["attached"])) == {0, 1, 2, 3}


def test_kernel_source_replays_the_reference_dense_and_reset_at_goal_rollouts():
    replay("dense", dense=True)
    replay("rag", dense=False)


def test_kernel_source_on_the_reference_closest_object_cases():
    emu = Emu(GOLD["attach_q0"])
    o, _, _ = emu.step(GOLD["attach_actions"])
    assert np.array_equal(emu.att, GOLD["attach_att"])
    assert np.array_equal(emu.q, GOLD["attach_q1"])
    assert np.array_equal(o, GOLD["attach_obs"])


def test_kernel_source_against_the_checker_on_a_random_batch():
    n, steps = 3001, 120
    rs = np.random.RandomState(17)
    q0 = np.concatenate([rs.uniform(-2, 2, (n, 2))] * 4, axis=1)
    q0[:, 2:] += rs.uniform(-0.6, 0.6, (n, 6))
    for dense in (False, True):
        emu, orc = Emu(q0, dense), tabletop3.Tabletop3Oracle(n, 1 << 40, dense=dense)
        orc.reset(init_qpos=q0)
        for t in range(steps):
            a = rs.uniform(-1.2, 1.2, (n, 3)).astype(np.float32)
            a[: n // 2, 2] = np.where(rs.uniform(size=n // 2) < 0.9, 1.0, -1.0)
            o, r, s = emu.step(a)
            o2, r2, _, s2 = orc.step(a)
            assert np.array_equal(o, o2) and np.array_equal(s, s2), t
            assert np.array_equal(emu.q, orc.qpos) and np.array_equal(emu.att, orc.att), t
            if dense:
                assert np.allclose(r, r2, rtol=2e-6, atol=2e-6)
            else:
                assert np.array_equal(r.astype(np.float64), r2)
        assert set(np.unique(orc.att)) == {0, 1, 2, 3}
