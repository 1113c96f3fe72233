#!/bin/bash
# debug build with device-side solver trace (never the shipped library): build to a side path, run, restore
mkdir -p gpurun_out
cp earl_benchmark_b200/libearl_b200.so /tmp/lib_keep.so
EARL_MJ_EXTRA_FLAGS=-DMJ_TRACE_DEVICE python -m earl_benchmark_b200.build --force > /dev/null
DIAG_ONLY_DRILL=1 python tools/diag_parity.py ${1:-3} > gpurun_out/trace.log 2>&1
cp /tmp/lib_keep.so earl_benchmark_b200/libearl_b200.so; touch earl_benchmark_b200/libearl_b200.so
grep -v "^state [0-9]*:" gpurun_out/trace.log | tail -${2:-60}
