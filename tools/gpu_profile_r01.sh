# Round-1 evidence run: tests, smoke, bench (both arms), ncu launch list + full captures of both step kernels.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 200 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_reference_r01.json
python bench.py --steps 2000 --warmup 100 2>&1 | tail -1 > gpurun_out/bench_r01.json
cat gpurun_out/bench_r01.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r01.csv python bench.py --profile --steps 100 --warmup 3 > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tabletop_step -s 20 -c 2 -o gpurun_out/prof_step_lsu_1M_r01 -f python bench.py --profile --steps 40 --warmup 3 > /dev/null 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:tabletop_step -s 20 -c 2 -o gpurun_out/prof_step_lsu_1M_warm_r01 -f python bench.py --profile --steps 40 --warmup 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tabletop_step -s 10 -c 2 -o gpurun_out/prof_step_tma_8M_r01 -f python bench.py --profile --num-envs 8388608 --steps 20 --warmup 3 > /dev/null 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu_info.csv
