#!/bin/bash
# small single-GPU experiments of round 2: inflate variants, chunk rounds, L2 fetch granularity vs DRAM traffic
mkdir -p gpurun_out
bash scripts/gpu_inflate_ab.sh "0" > /dev/null 2>&1; cat gpurun_out/inflate_ab.log
for r in 2 3 4; do
  FASTF_CHUNK_ROUNDS=$r python bench.py --base-reads 16000000 --reads 128000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-hw-extra --freq-reads 0 --check-reads 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('chunk rounds $r: %.1f Mreads/s, ms/step %.1f, inflate %.1f ms, chunks %d' % (d['value']/1e6, d['ms_per_step'], d['stages']['inflate']['ms'], d['config']['counters']['n_chunks']))"
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for g in default 32 64 128; do
  if [ $g = default ]; then unset FASTF_L2_FETCH; else export FASTF_L2_FETCH=$g; fi
  python scripts/inflate_ab.py --reps 1 > /dev/null 2>&1 && ncu --metrics $M --clock-control none -k regex:inflate_tps -c 1 --csv --log-file gpurun_out/ncu_l2fetch_$g.csv python scripts/inflate_ab.py --reps 1 > /dev/null 2>&1
  echo "L2 fetch granularity $g: $(grep -v '^==' gpurun_out/ncu_l2fetch_$g.csv | python -c "
import csv,sys
rows=[r for r in csv.reader(sys.stdin) if len(r)>10]
h=rows[0]; print(' '.join('%s=%s%s' % (r[h.index('Metric Name')].split('.')[0], r[h.index('Metric Value')], r[h.index('Metric Unit')]) for r in rows[1:]))")"
done
