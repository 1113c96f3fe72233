"""Device engine of the kitchen capacity set (include/earl_mj_kitchen_b200.h) against the fp64 checker: the lane-parallel
code paths of the joint-equality / friction-loss / pyramidal rows and of the capsule collisions on the B200."""
import numpy as np
import pytest
import torch

from earl_benchmark_b200.kitchen_engine import KitchenEngine
from oracle import kitchen_logic as KL
from oracle.engine import KitchenOracle

pytestmark = pytest.mark.gpu
NV = 23


def _states(count):
    """(qpos, qvel, warm, mocap, ctrl) along a scripted checker rollout that reaches into the cabinets."""
    from earl_benchmark_b200.mjcf.compile import Model
    from earl_benchmark_b200.kitchen_engine import MODEL_PATH
    k = KitchenOracle(Model.load(MODEL_PATH))
    k.seed(3)
    np.random.seed(2)
    k.reset()
    e = k.e
    target = e.site_xpos("slide_site").copy()
    out = []
    for t in range(100):
        d = target - e.site_xpos("end_effector")
        if t >= 70:
            d = np.array([0.5, 0.0, 0.0])
        a = np.zeros(9)
        a[:3] = np.clip(d * 10, -1, 1) * 0.5
        mocap, ctrl = k.logic.control(a, e.mocap_pos.copy())
        e.mocap_pos[:], e.ctrl[:] = mocap, ctrl
        if t % max(1, 100 // count) == 0 and len(out) < count:
            out.append((e.qpos.copy(), e.qvel.copy(), e.arr("qacc_warmstart", (32,))[:NV].copy(), e.mocap_pos.copy(), e.ctrl.copy()))
        e.step(KL.FRAME_SKIP)
        k.logic.observe(e.qpos)
    return k, out


def test_substep_parity_and_batch_independence():
    k, states = _states(50)
    e = k.e
    eng = KitchenEngine("cuda:0")
    dev = eng.device
    f32 = lambda a: torch.tensor(np.stack(a), dtype=torch.float32, device=dev)  # noqa: E731
    q, v, w = f32([s[0] for s in states]), f32([s[1] for s in states]), f32([s[2] for s in states])
    mp = torch.tensor(np.stack([s[3] for s in states]), dtype=torch.float64, device=dev)
    c = f32([s[4] for s in states])
    q1, v1, w1 = q.clone(), v.clone(), w.clone()
    info = eng.substeps(q1, v1, w1, mp, c, nsub=4).cpu().numpy()
    # no failure, nothing dropped; the states with many six-dimensional finger contacts outgrow the 112 rows of the primary
    # set and went through the 544-row set (bit 4) -- the comparison below covers them like any other state
    assert np.all((info[:, 3] & 15) == 0) and np.any(info[:, 3] & 16) and not np.all(info[:, 3] & 16)
    dq, dv, con = [], [], 0
    for i, (qq, vv, ww, mm, cc) in enumerate(states):
        e.reset()
        e.qpos[:], e.qvel[:], e.mocap_pos[:], e.ctrl[:] = qq.astype(np.float32), vv.astype(np.float32), mm, cc
        e.arr("qacc_warmstart", (32,))[:NV] = ww.astype(np.float32)
        e.step(4)
        dq.append(np.abs(q1[i].cpu().numpy() - e.qpos).max())
        dv.append(np.abs(v1[i].cpu().numpy() - e.qvel).max())
        assert info[i, 1] == e.ncon and abs(info[i, 0] - e.nefc) <= 2
        con = max(con, e.ncon)
    dq, dv = np.array(dq), np.array(dv)
    print("kitchen engine vs checker, 4 substeps: dq max %.2e, dv median %.2e max %.2e, max contacts %d" % (dq.max(), np.median(dv), dv.max(), con))
    assert con >= 5
    assert dq.max() < 1e-4 and np.median(dv) < 5e-5 and np.percentile(dv, 90) < 1e-3 and dv.max() < 0.5
    # results do not depend on the batch an environment sits in
    idx = torch.tensor([7, 3, 41], device=dev)
    q2, v2, w2 = q[idx].clone(), v[idx].clone(), w[idx].clone()
    eng.substeps(q2, v2, w2, mp[idx].clone(), c[idx].clone(), nsub=4)
    assert torch.equal(q2, q1[idx]) and torch.equal(v2, v1[idx]) and torch.equal(w2, w1[idx])
