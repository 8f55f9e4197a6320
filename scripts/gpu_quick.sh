#!/bin/bash
# quick A/B: stage timings of one 8M-read job per inflate lane width (parity is checked by the gpu tests, run first)
mkdir -p gpurun_out
./fastf_b200/_build/de_probe > gpurun_out/de_probe.txt 2>&1; cat gpurun_out/de_probe.txt
python -m pytest tests -m gpu -x -q -k "inflate or streaming or config1" 2>&1 | tail -3
for lanes in 0 32; do
python bench.py --reads 64000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --lanes $lanes ${EXTRA} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('lanes $lanes', 'Mreads/s %.1f'%(d['value']/1e6), 'ms/step %.1f'%d['ms_per_step'], {k:v['ms'] for k,v in d['stages'].items()}, 'chunks', d['config']['counters']['n_chunks'])"
done
