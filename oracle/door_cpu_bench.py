"""CPU baseline worker for bench.py: one process = one fp64 checker instance stepping the Sawyer door (or peg) env with random
actions.  TEST / BENCH INFRASTRUCTURE ONLY (bench.py's cpu_baseline leg).  Prints one JSON line:
{"seconds": wall time of the stepping loop, "steps": n, "flops_per_env_step": checker flop counter / n}."""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main(seed, steps, task="sawyer_door"):
    from earl_benchmark_b200.mjcf.compile import Model
    from oracle.engine import KitchenOracle, SawyerDoorOracle, SawyerPegOracle
    model = Model.load(os.path.join(REPO, "earl_benchmark_b200", "models", task + ".npz"))
    rs = np.random.RandomState(seed)
    if task == "kitchen":
        o = KitchenOracle(model)
        o.seed(seed)
        np.random.seed(seed)
        o.reset()
        acts = rs.uniform(-1, 1, (steps, 9))
        for a in acts[:3]:
            o.step(a)
        f0 = o.e.flops
        t0 = time.perf_counter()
        for a in acts:
            o.step(a)
        print(json.dumps({"seconds": time.perf_counter() - t0, "steps": steps, "flops_per_env_step": (o.e.flops - f0) / steps}))
        return
    if task == "sawyer_door":
        o = SawyerDoorOracle(model)
        o.reset(door_angle=-np.pi / 3 + rs.uniform(0, np.pi / 20))
    else:
        o = SawyerPegOracle(model)
        o.reset()
    acts = rs.uniform(-1, 1, (steps, 4))
    for a in acts[:10]:
        o.step(a)
    f0 = o.e.flops
    t0 = time.perf_counter()
    for a in acts:
        o.step(a)
    print(json.dumps({"seconds": time.perf_counter() - t0, "steps": steps, "flops_per_env_step": (o.e.flops - f0) / steps}))


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]), sys.argv[3] if len(sys.argv) > 3 else "sawyer_door")
