set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01.csv python bench.py --profile --steps 40 --warmup 3 > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tabletop_step -s 10 -c 3 -o gpurun_out/prof_step_r01 -f python bench.py --profile --steps 40 --warmup 3 > gpurun_out/prof_bench.log 2>&1
for n in 65536 262144 1048576 4194304 8388608 16777216; do python bench.py --num-envs $n --steps 1000 --warmup 50 --no-cpu-baseline --e2e-steps 20 2>&1 | tail -1 > gpurun_out/sweep_$n.json; done
cat gpurun_out/sweep_*.json | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['config']['envs_per_gpu'], '%.3e'%d['value'], 'us/step %.2f'%(d['ms_per_step']*1e3), 'frac %.3f'%d['roofline']['frac'], 'e2e %.3e'%d['e2e']['value'], d['clocks'])
"
