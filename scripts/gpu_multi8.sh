#!/bin/bash
# 8-GPU box (gpurun --gpus 8): concurrent H2D probe (what bounds the e2e leg at 8 ranks), the bench line at 8 ranks with the depth sweep
# (BASELINE configs[4]), the C one-process driver on 8 GPUs against the reference CLI.
# usage: gpu_multi8.sh [N=8] [tag=r02] [bench reads per GPU=64000000] [parity reads=16000000]
N=${1:-8}; TAG=${2:-r02}; BREADS=${3:-64000000}; PREADS=${4:-16000000}
mkdir -p gpurun_out
( time python scripts/h2d_probe_multi.py ) > gpurun_out/${TAG}_h2d_probe_n$N.txt 2>&1; tail -6 gpurun_out/${TAG}_h2d_probe_n$N.txt
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --reads $BREADS --steps 2 --warmup 1 --depth-sweep --no-hw-extra --no-cpu-baseline --freq-reads 0 ) \
   > gpurun_out/${TAG}_bench_n${N}_depth_sweep.json 2> gpurun_out/${TAG}_bench_n$N.err; tail -3 gpurun_out/${TAG}_bench_n$N.err; head -c 800 gpurun_out/${TAG}_bench_n${N}_depth_sweep.json
[ "$PREADS" -gt 0 ] && bash scripts/gpu_parity_scale.sh $PREADS $N | tail -22
