#!/bin/bash
# bench line + ncu launch list + one full capture of the dominant kernel (each ncu run only after the same command exited 0 plain)
mkdir -p gpurun_out
SMALL="python bench.py --reads 64000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
$SMALL > gpurun_out/small.json 2> gpurun_out/small.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_small.log 2>&1
echo "ncu list rc=$?"
$SMALL > gpurun_out/small2.json 2> gpurun_out/small2.err &&
ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNEL:-fastf_bgzf_inflate}" -s 2 -c 2 -o gpurun_out/prof_${NCU_TAG:-inflate} -f $SMALL > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
