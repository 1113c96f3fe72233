#!/usr/bin/env python
"""Diagnostic: per-state one-step parity of the device engine against the checker (same states as
tests/test_sawyer_door_gpu.py::test_one_step_parity_from_identical_states), worst states first."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
from earl_benchmark_b200.envs import sawyer_door  # noqa: E402
from earl_benchmark_b200.mjcf.compile import Model  # noqa: E402
from oracle.engine import SawyerDoorOracle  # noqa: E402
import test_sawyer_door_gpu as T  # noqa: E402

o = SawyerDoorOracle(Model.load(sawyer_door.MODEL_PATH))
n = 96
states = T._reference_states(o, n, seed=int(sys.argv[1]) if len(sys.argv) > 1 else 3)
rs = np.random.RandomState(4)
actions = rs.uniform(-1.2, 1.2, (n, 4)).astype(np.float32)
sub = int(sys.argv[2]) if len(sys.argv) > 2 else 0
env = sawyer_door.SawyerDoorV2(num_envs=n, device="cuda:0")
env.reset()
env.set_state(qpos=np.stack([s[0] for s in states]), qvel=np.stack([s[1] for s in states]),
              qacc_warmstart=np.stack([s[2] for s in states]), mocap_pos=np.stack([s[3] for s in states]))
w0 = env.work_counters()
env.step(torch.from_numpy(actions).cuda())
got = env.get_state()
w1 = env.work_counters()
print({k: w1[k] - w0[k] for k in w1})
e = o.e
rows = []
for i, (q, v, w, mp) in enumerate(states):
    e.reset()
    e.qpos[:], e.qvel[:], e.mocap_pos[:] = q, v, mp
    e.arr("qacc_warmstart", (32,))[:e.nv] = w
    e.forward()
    n0 = e.ncon
    it = 0
    a = np.clip(actions[i].astype(np.float64), -1, 1)
    e.mocap_pos[:] = np.clip(e.mocap_pos + a[:3] * 0.01, o.MOCAP_LOW, o.MOCAP_HIGH)
    e.mocap_quat[:] = [1, 0, 1, 0]
    e.ctrl[:] = [a[3], -a[3]]
    ncs, its = [], []
    for s in range(5):
        e.step(1)
        ncs.append(e.ncon)
        its.append(e.solver_iter)
    dq = np.abs(got["qpos"][i] - e.qpos)
    dv = np.abs(got["qvel"][i] - e.qvel)
    rows.append((dv.max(), dq.max(), i, n0, ncs, its, int(dv.argmax())))
rows.sort(key=lambda r: -r[0])
for r in rows[:8]:
    print("state %d: dv %.2e (dof %d) dq %.2e  ncon0 %d  checker ncon/substep %s iters %s" % (r[2], r[0], r[6], r[1], r[3], r[4], r[5]))

# the worst state alone: device (one substep per launch) vs one-lane host build of the same source vs checker
from host_emulation.emu import Emu, door_task  # noqa: E402
m = Model.load(sawyer_door.MODEL_PATH)
em = Emu(m, door_task(m))
_orig = sawyer_door.task_spec


def _one_substep(model, *a, **kw):
    t = _orig(model, *a, **kw)
    t.frame_skip = 1
    return t


sawyer_door.task_spec = _one_substep
for r in rows[:1 if os.environ.get('DIAG_ONLY_DRILL') else 2]:
    i = r[2]
    q, v, w, mp = states[i]
    a = np.clip(actions[i].astype(np.float64), -1, 1)
    mp2 = np.clip(mp + a[:3] * 0.01, o.MOCAP_LOW, o.MOCAP_HIGH)
    env1 = sawyer_door.SawyerDoorV2(num_envs=1, device="cuda:0")
    env1.reset()
    env1.set_state(qpos=q[None], qvel=v[None], qacc_warmstart=w[None], mocap_pos=mp[None])
    em.set_state(q, v, w, mp2, ctrl=(a[3], -a[3]))
    e.reset()
    e.qpos[:], e.qvel[:], e.mocap_pos[:] = q, v, mp2
    e.arr("qacc_warmstart", (32,))[:e.nv] = w
    e.mocap_quat[:] = [1, 0, 1, 0]
    e.ctrl[:] = [a[3], -a[3]]
    print("state", i)
    for s in range(5):
        act = actions[i:i + 1].copy()
        if s > 0:
            act[0, :3] = 0
        c0 = env1.work_counters()
        print("  -- substep", s, flush=True)
        env1.step(torch.from_numpy(act).cuda())
        torch.cuda.synchronize()
        c1 = env1.work_counters()
        g1 = env1.get_state()
        e.step(1)
        em.substeps(1)
        q2, v2, w2, _ = em.get_state()
        if os.environ.get('DIAG_ONLY_DRILL'):
            hd, hp, hf, hg = em.contacts()
            for c in range(len(hd)):
                print("  [host-build] con %d g %s dist %.9g pos %s n %s" % (c, hg[c], hd[c], hp[c], hf[c][:3]))
            nc = e.ncon
            cd, cp, cf = e.arr("con_dist", (64,))[:nc], e.arr("con_pos", (64, 3))[:nc], e.arr("con_frame", (64, 9))[:nc]
            for c in range(nc):
                print("  [checker] con %d dist %.9g pos %s n %s" % (c, cd[c], cp[c], cf[c][:3]))
        print("  substep", s, "iters checker/host-build/device", e.solver_iter, em.info("iter"), c1["newton_iterations"] - c0["newton_iterations"],
              "rows", e.nefc, c1["constraint_rows"] - c0["constraint_rows"],
              "| dv host-build vs checker %.2e  device vs checker %.2e  dwarm device vs checker %.2e" % (
                  np.abs(v2 - e.qvel).max(), np.abs(g1["qvel"][0] - e.qvel).max(),
                  np.abs(g1["qacc_warmstart"][0] - e.arr("qacc_warmstart", (32,))[:e.nv]).max()))
