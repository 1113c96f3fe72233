#!/bin/bash
# round-1 GPU pass for the Sawyer door engine: tests, throughput sweep, launch list, one full ncu capture
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu_info_door.csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_door.py --envs 1024 4096 16384 65536 --steps 30 --warmup 5 > gpurun_out/bench_door.log 2>&1
cat gpurun_out/bench_door.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_door.csv \
  python tools/bench_door.py --envs 16384 --steps 10 --warmup 2 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mj_step_kernel -s 3 -c 1 -o gpurun_out/prof_door_step \
  python tools/bench_door.py --envs 16384 --steps 6 --warmup 2 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
