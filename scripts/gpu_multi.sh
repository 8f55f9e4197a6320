#!/bin/bash
# Multi-GPU validation (gpurun --gpus N): NCCL parity tests, the C one-process driver against the reference CLI, the bench line at N ranks.
# usage: gpu_multi.sh <N> [tag=r02] [parity reads=16000000] [bench reads per GPU=64000000] [extra bench args]
N=${1:-2}; TAG=${2:-r02}; PREADS=${3:-16000000}; BREADS=${4:-64000000}; shift 4
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > gpurun_out/${TAG}_gpus_n$N.txt
( time python -m pytest tests -m gpu -x -q -k "two_gpus or sharded_driver_nccl" ) > gpurun_out/${TAG}_pytest_gpu_n$N.log 2>&1; tail -4 gpurun_out/${TAG}_pytest_gpu_n$N.log
[ "$PREADS" -gt 0 ] && bash scripts/gpu_parity_scale.sh $PREADS $N | tail -14
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --reads $BREADS --steps 2 --warmup 1 --no-hw-extra --no-cpu-baseline --freq-reads 0 "$@" ) \
   > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; tail -3 gpurun_out/${TAG}_bench_n$N.err; head -c 1200 gpurun_out/${TAG}_bench_n$N.json
