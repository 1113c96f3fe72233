# Round-2 evidence run (tabletop): ncu launch list of the bench command + full captures of the step kernels.
mkdir -p gpurun_out/r02
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02/launches_r02.csv python bench.py --profile --steps 100 --warmup 3 > gpurun_out/r02/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tabletop_step -s 20 -c 2 -o gpurun_out/r02/prof_step_lsu_1M_r02 -f python bench.py --profile --steps 40 --warmup 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tabletop_step -s 10 -c 2 -o gpurun_out/r02/prof_step_tile_8M_r02 -f python bench.py --profile --num-envs 8388608 --steps 20 --warmup 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tabletop_step -s 10 -c 2 -o gpurun_out/r02/prof_step_tile_4M_r02 -f python bench.py --profile --num-envs 4194304 --steps 20 --warmup 3 > /dev/null 2>&1
ls -la gpurun_out/r02
