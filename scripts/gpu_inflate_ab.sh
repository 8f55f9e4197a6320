#!/bin/bash
# A/B of the inflate kernel alone over library variants (scripts/build_variants.sh) on one chunk-sized BGZF image
# usage: gpu_inflate_ab.sh [ncu-variant ...]   -- every variant is timed; the named ones also get an ncu counter pass
mkdir -p gpurun_out
LOG=gpurun_out/inflate_ab.log
: > $LOG
python scripts/inflate_ab.py --reps 1 2>gpurun_out/ab_err.txt >> $LOG || tail -3 gpurun_out/ab_err.txt >> $LOG  # generates + caches the image, default in-tree build
for lib in fastf_b200/_build/variants/*.so; do
  FASTF_GPU_LIB=$PWD/$lib python scripts/inflate_ab.py 2>gpurun_out/ab_err.txt >> $LOG || { echo "FAILED $lib" >> $LOG; tail -2 gpurun_out/ab_err.txt >> $LOG; }
done
for g in 32 64 128; do
  echo "FASTF_L2_FETCH=$g" >> $LOG
  FASTF_L2_FETCH=$g python scripts/inflate_ab.py 2>gpurun_out/ab_err.txt >> $LOG || tail -2 gpurun_out/ab_err.txt >> $LOG
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_op_write_lookup_miss.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
for v in "$@"; do
  lib=$PWD/fastf_b200/_build/variants/$v.so
  [ "$v" = default ] && lib=$PWD/fastf_b200/_build/libfastf_gpu.so
  FASTF_AB_NOCHECK=1 FASTF_GPU_LIB=$lib python scripts/inflate_ab.py --reps 2 > gpurun_out/plain_$v.log 2>&1 &&
  FASTF_AB_NOCHECK=1 FASTF_GPU_LIB=$lib ncu --metrics $M --clock-control none -k regex:inflate --csv --log-file gpurun_out/ncu_ab_$v.csv python scripts/inflate_ab.py --reps 2 > gpurun_out/ncu_ab_$v.log 2>&1
  echo "ncu $v rc=$?" >> $LOG
done
cat $LOG
