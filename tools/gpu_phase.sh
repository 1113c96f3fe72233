#!/bin/bash
# profiling build with per-phase cycle counters (never the shipped library): build to a side path, run, restore
set -x
mkdir -p gpurun_out
cp earl_benchmark_b200/libearl_b200.so /tmp/lib_keep.so
EARL_MJ_EXTRA_FLAGS=-DMJ_PHASE_TIMING python -m earl_benchmark_b200.build --force > /dev/null
python tools/bench_door.py --envs 16384 --steps 100 --warmup ${PHASE_WARMUP:-5} > gpurun_out/phase.log 2>&1
cp /tmp/lib_keep.so earl_benchmark_b200/libearl_b200.so; touch earl_benchmark_b200/libearl_b200.so
cat gpurun_out/phase.log
