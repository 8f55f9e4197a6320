mkdir -p gpurun_out
python bench.py --workload freq --reads 100000000 --steps 2 --warmup 1 > gpurun_out/freq_bench.json 2> gpurun_out/freq_bench.err; echo rc=$?; tail -2 gpurun_out/freq_bench.err; cut -c1-1800 gpurun_out/freq_bench.json
SMALL="python bench.py --workload freq --reads 16000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$SMALL > gpurun_out/freq_small.json 2>/dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/freq_launches.csv $SMALL > gpurun_out/freq_ncu.log 2>&1; echo "ncu rc=$?"
