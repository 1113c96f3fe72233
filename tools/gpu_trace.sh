#!/bin/bash
# debug build with device-side solver trace (never the shipped library): build to a side path, run, restore
mkdir -p gpurun_out
cp earl_benchmark_b200/libearl_b200.so /tmp/lib_keep.so
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden --fmad=false -c -o /tmp/a.o earl_benchmark_b200/csrc/earl_b200.cu
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -DMJ_TRACE_DEVICE -c -o /tmp/b.o earl_benchmark_b200/csrc/earl_mj.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o earl_benchmark_b200/libearl_b200.so /tmp/a.o /tmp/b.o
DIAG_ONLY_DRILL=1 python tools/diag_parity.py ${1:-3} > gpurun_out/trace.log 2>&1
cp /tmp/lib_keep.so earl_benchmark_b200/libearl_b200.so
grep -v "^state [0-9]*:" gpurun_out/trace.log | tail -${2:-60}
