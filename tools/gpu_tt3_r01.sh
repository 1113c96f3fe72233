#!/bin/bash
# Round-1 GPU pass for the three-object tabletop: tests, smoke, bench line, launch list and one ncu capture.
set -x
export PYTHONPATH=$PWD
mkdir -p gpurun_out/tt3
python -m pytest tests/test_tabletop3_gpu.py -x -q > gpurun_out/tt3/pytest_tt3.log 2>&1; tail -15 gpurun_out/tt3/pytest_tt3.log
python -m pytest tests -q -m gpu > gpurun_out/tt3/pytest_gpu_all.log 2>&1; tail -5 gpurun_out/tt3/pytest_gpu_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/tt3/smoke.log 2>&1; tail -6 gpurun_out/tt3/smoke.log
python bench.py > gpurun_out/tt3/bench.json 2> gpurun_out/tt3/bench.err; tail -c 1500 gpurun_out/tt3/bench.json
cat > /tmp/prof3.py <<'P'
import torch
from earl_benchmark_b200.envs.tabletop_manipulation_3obj import TabletopManipulation
n = 1 << 22
env = TabletopManipulation(reward_type="sparse", num_envs=n, device="cuda:0")
env.reset()
a = torch.rand((4, n, 3), device="cuda") * 2 - 1
obs = torch.empty((2, n, 20), device="cuda"); rew = torch.empty((2, n), device="cuda"); done = torch.empty((2, n), dtype=torch.uint8, device="cuda")
env.rollout_into(a, 12, obs, rew, done)
torch.cuda.synchronize()
P
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/tt3/launches_tt3.csv python /tmp/prof3.py > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tt3_step_kernel -s 8 -c 1 -o gpurun_out/tt3/prof_tt3_step_4M python /tmp/prof3.py > gpurun_out/tt3/ncu.log 2>&1
ncu -i gpurun_out/tt3/prof_tt3_step_4M.ncu-rep --page raw --csv > gpurun_out/tt3/prof_tt3_step_4M.raw.csv 2>/dev/null
ncu -i gpurun_out/tt3/prof_tt3_step_4M.ncu-rep --page details --csv > gpurun_out/tt3/prof_tt3_step_4M.details.csv 2>/dev/null
ls -la gpurun_out/tt3
