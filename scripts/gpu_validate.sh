#!/bin/bash
# Round validation on one B200: GPU test-suite, the driver's bench line, parity at 64 M distinct reads against the reference CLI,
# launch list + full ncu capture of the dominant kernel.  Logs go to gpurun_out/ (copied to profiles/ afterwards).
# usage: gpu_validate.sh [tag=r02] [parity reads=64000000]
TAG=${1:-r02}; PREADS=${2:-64000000}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -4 gpurun_out/${TAG}_pytest_gpu.log
( time python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err; head -c 1500 gpurun_out/${TAG}_bench.json
( time python bench.py --impl reference --steps 1 --warmup 0 ) > gpurun_out/${TAG}_bench_reference_arm.json 2>> gpurun_out/${TAG}_bench.err
[ "$PREADS" -gt 0 ] && bash scripts/gpu_parity_scale.sh $PREADS 1 | tail -14
CMD="python bench.py --reads 64000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-hw-extra --freq-reads 0 --check-reads 0"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/${TAG}_plain2.json 2> gpurun_out/${TAG}_plain2.err &&
ncu --set full --clock-control none --import-source on -k regex:"inflate_tps|bam_parse|crc32" -s 3 -c 3 -f -o gpurun_out/prof_${TAG}_top $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"
