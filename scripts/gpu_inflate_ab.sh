#!/bin/bash
# A/B of the inflate kernel alone over library variants (scripts/build_variants.sh); every variant gets an image of exactly two rounds
# of its own stream count (scripts/inflate_ab.py).
# usage: gpu_inflate_ab.sh "<lanes values for the default-shape builds>" [ncu-full-variant[:lanes] ...]
mkdir -p gpurun_out
LOG=gpurun_out/inflate_ab.log
: > $LOG
LANES="${1:-0}"; shift
python scripts/inflate_ab.py --reps 1 2>gpurun_out/ab_err.txt >> $LOG || tail -3 gpurun_out/ab_err.txt >> $LOG   # generates + caches the image
for lib in fastf_b200/_build/variants/*.so; do
  NOCRC=""; case "$lib" in *nocopy*|*_l1*) NOCRC=1;; esac
  LL="0"; case "$(basename $lib)" in zz*) LL="$LANES";; esac
  for l in $LL; do
    FASTF_AB_NOCRC=$NOCRC FASTF_GPU_LIB=$PWD/$lib timeout 200 python scripts/inflate_ab.py --lanes $l 2>gpurun_out/ab_err.txt >> $LOG || { echo "FAILED $lib lanes $l" >> $LOG; tail -2 gpurun_out/ab_err.txt >> $LOG; }
  done
done
for spec in "$@"; do
  v="${spec%%:*}"; l="${spec#*:}"; [ "$l" = "$spec" ] && l=0
  lib=$PWD/fastf_b200/_build/variants/$v.so
  [ "$v" = default ] && lib=$PWD/fastf_b200/_build/libfastf_gpu.so
  FASTF_GPU_LIB=$lib python scripts/inflate_ab.py --lanes $l --reps 1 > gpurun_out/plain_$v.log 2>&1 &&
  FASTF_GPU_LIB=$lib timeout 400 ncu --set full --clock-control none --import-source on -k regex:inflate_tps -c 1 -f -o gpurun_out/prof_r02_$v python scripts/inflate_ab.py --lanes $l --reps 1 > gpurun_out/ncu_ab_$v.log 2>&1
  echo "ncu $v rc=$?" >> $LOG
done
cat $LOG
