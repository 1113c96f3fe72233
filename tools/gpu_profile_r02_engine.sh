#!/bin/bash
# round-2 evidence for the Sawyer engine: throughput sweep, launch list (step / redo / order kernels), full ncu captures
mkdir -p gpurun_out/r02
for t in sawyer_door sawyer_peg; do
  timeout 600 python tools/bench_door.py --envs 4096 16384 65536 --steps 100 --warmup 100 --task $t > gpurun_out/r02/bench_${t}_sweep_r02.jsonl 2>&1
done
cat gpurun_out/r02/bench_sawyer_door_sweep_r02.jsonl gpurun_out/r02/bench_sawyer_peg_sweep_r02.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 60 --csv --log-file gpurun_out/r02/launches_door_r02.csv \
  python tools/bench_door.py --envs 16384 --steps 20 --warmup 100 > gpurun_out/r02/ncu_launch_door.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mj_step_kernel -s 105 -c 1 -o gpurun_out/r02/prof_door_steady_16k_r02 -f \
  python tools/bench_door.py --envs 16384 --steps 8 --warmup 100 > gpurun_out/r02/ncu_full_door.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mj_redo_kernel -s 105 -c 1 -o gpurun_out/r02/prof_redo_16k_r02 -f \
  python tools/bench_door.py --envs 16384 --steps 8 --warmup 100 --task sawyer_peg > gpurun_out/r02/ncu_full_redo.log 2>&1
ls -la gpurun_out/r02
