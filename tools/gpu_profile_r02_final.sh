#!/bin/bash
# round-2 final evidence: the driver's bench command, the kitchen launch list and one full ncu capture of the kitchen step kernel
mkdir -p gpurun_out/r02
python bench.py --steps 20 --warmup 5 > gpurun_out/r02/bench_r02_1gpu.json 2> gpurun_out/r02/bench_r02_1gpu.err
tail -c 600 gpurun_out/r02/bench_r02_1gpu.err
python tools/bench_kitchen.py --envs 4736 14208 --steps 8 --warmup 30 > gpurun_out/r02/bench_kitchen_r02.jsonl 2>&1
cat gpurun_out/r02/bench_kitchen_r02.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/r02/launches_kitchen_r02.csv \
  python tools/bench_kitchen.py --envs 4736 --steps 8 --warmup 30 > gpurun_out/r02/ncu_launch_kitchen.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:mjk_task_kernel -s 32 -c 1 -o gpurun_out/r02/prof_kitchen_steady_4736_r02 -f \
  python tools/bench_kitchen.py --envs 4736 --steps 4 --warmup 32 > gpurun_out/r02/ncu_full_kitchen.log 2>&1
tail -3 gpurun_out/r02/ncu_full_kitchen.log
ls -la gpurun_out/r02
