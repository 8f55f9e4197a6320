#!/bin/bash
# one full ncu capture of selected kernels on a short bench command (plain run first, per the profiling recipe)
# usage: gpu_ncu.sh <kernel-regex> <tag> [extra bench args]
mkdir -p gpurun_out
K="$1"; TAG="$2"; shift 2
CMD="python bench.py --reads 16000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-hw-extra $*"
$CMD > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.err &&
ncu --set full --clock-control none --import-source on -k regex:"$K" -s 4 -c 4 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$TAG.log
