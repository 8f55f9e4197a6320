#!/bin/bash
# A/B of thread-per-stream inflate shapes: parity subset first, then stage timings of a 64M-read job per shape
# usage: gpu_ab.sh "<lanes values>" [extra bench args]
mkdir -p gpurun_out
LANES="${1:-1 2 3 4}"; shift
python -m pytest tests -m gpu -x -q -k "inflate or streaming or config1" 2>&1 | tail -3
for lanes in $LANES; do
python bench.py --reads 64000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-hw-extra --lanes $lanes "$@" 2>gpurun_out/ab_$lanes.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('lanes $lanes', 'Mreads/s %.1f'%(d['value']/1e6), 'ms/step %.1f'%d['ms_per_step'], {k:v['ms'] for k,v in d['stages'].items()}, 'chunks', d['config']['counters']['n_chunks'])"
done
