#!/bin/bash
# final build of round 2: SURVEY 8(d) config 3 / 4 / 5 sweeps (door, peg: 4,096 / 16,384 / 65,536 envs per GPU; kitchen:
# 1,024 / 4,096 / 16,384 and the kernel's natural batch sizes 4,736 / 14,208) and one full ncu capture of the step kernel on
# the peg (the door capture is prof_door_steady_16k_r02)
O=gpurun_out/r02f
mkdir -p $O
for t in sawyer_door sawyer_peg; do
  timeout 200 python tools/bench_door.py --envs 4096 16384 65536 --steps 100 --warmup 100 --task $t > $O/bench_${t}_sweep_final.jsonl 2>$O/bench_${t}.err
  cat $O/bench_${t}_sweep_final.jsonl
done
timeout 240 python tools/bench_kitchen.py --envs 1024 4096 4736 14208 16384 --steps 8 --warmup 30 > $O/bench_kitchen_sweep_final.jsonl 2>$O/bench_kitchen.err
cat $O/bench_kitchen_sweep_final.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mj_step_kernel -s 105 -c 1 -o $O/prof_peg_steady_16k_final -f \
  python tools/bench_door.py --envs 16384 --steps 8 --warmup 100 --task sawyer_peg > $O/ncu_full_peg.log 2>&1
tail -2 $O/ncu_full_peg.log
ls -la $O
