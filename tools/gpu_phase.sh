#!/bin/bash
# profiling build with per-phase cycle counters (never the shipped library): build to a side path, run, restore
set -x
mkdir -p gpurun_out
cp earl_benchmark_b200/libearl_b200.so /tmp/lib_keep.so
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden --fmad=false -c -o /tmp/a.o earl_benchmark_b200/csrc/earl_b200.cu
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -DMJ_PHASE_TIMING -c -o /tmp/b.o earl_benchmark_b200/csrc/earl_mj.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o earl_benchmark_b200/libearl_b200.so /tmp/a.o /tmp/b.o
python tools/bench_door.py --envs 16384 --steps 100 --warmup ${PHASE_WARMUP:-5} > gpurun_out/phase.log 2>&1
cp /tmp/lib_keep.so earl_benchmark_b200/libearl_b200.so
cat gpurun_out/phase.log
