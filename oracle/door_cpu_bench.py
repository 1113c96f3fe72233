"""CPU baseline worker for bench.py: one process = one fp64 checker instance stepping the Sawyer door env with random
actions.  TEST / BENCH INFRASTRUCTURE ONLY (bench.py's cpu_baseline leg).  Prints one JSON line:
{"seconds": wall time of the stepping loop, "steps": n, "flops_per_env_step": checker flop counter / n}."""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main(seed, steps):
    from earl_benchmark_b200.mjcf.compile import Model
    from oracle.engine import SawyerDoorOracle
    o = SawyerDoorOracle(Model.load(os.path.join(REPO, "earl_benchmark_b200", "models", "sawyer_door.npz")))
    rs = np.random.RandomState(seed)
    o.reset(door_angle=-np.pi / 3 + rs.uniform(0, np.pi / 20))
    acts = rs.uniform(-1, 1, (steps, 4))
    for a in acts[:10]:
        o.step(a)
    f0 = o.e.flops
    t0 = time.perf_counter()
    for a in acts:
        o.step(a)
    print(json.dumps({"seconds": time.perf_counter() - t0, "steps": steps, "flops_per_env_step": (o.e.flops - f0) / steps}))


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]))
