#!/bin/bash
# e2e (host-buffer) path of the tabletop step: chunk count and per-chunk vs tail copies of the small outputs
export PYTHONPATH=$PWD
mkdir -p gpurun_out/e2e
cat > /tmp/e2e.py <<'P'
import os, sys, time, torch
import earl_benchmark_b200 as eb
n = 1 << 20
tr, _ = eb.EARLEnvs("tabletop_manipulation", reward_type="sparse", num_envs=n, device="cuda:0", seed=0, train_horizon=200000).get_envs()
tr.reset()
ha = [torch.rand((n, 3)).mul_(2).sub_(1).pin_memory() for _ in range(4)]
for t in range(20): tr.step(ha[t % 4])
best = 0
for rep in range(3):
    t0 = time.perf_counter()
    for t in range(200): tr.step(ha[t % 4])
    best = max(best, n * 200 / (time.perf_counter() - t0))
print(os.environ.get("EARL_TT_HOST_TAIL", "-"), os.environ.get("EARL_TT_HOST_CHUNKS", "-"), f"{best:.4e}", flush=True)
P
for cfg in "0 8" "1 8" "1 4" "1 16" "1 2" "0 4" "0 16" "1 1"; do set -- $cfg; EARL_TT_HOST_TAIL=$1 EARL_TT_HOST_CHUNKS=$2 python /tmp/e2e.py 2>&1 | tail -1 | tee -a gpurun_out/e2e/sweep.txt; done
