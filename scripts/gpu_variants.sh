#!/bin/bash
# A/B of library variants built by build_variants.sh: stage timings of a 64M-read job per variant and lane shape
# usage: gpu_variants.sh "<lanes values>" [extra bench args]
mkdir -p gpurun_out
LANES="${1:-1 2}"; shift
for lib in fastf_b200/_build/variants/*.so; do
for lanes in $LANES; do
FASTF_GPU_LIB=$PWD/$lib python bench.py --reads 64000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-hw-extra --lanes $lanes "$@" 2>gpurun_out/var.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$(basename $lib) lanes $lanes', 'Mreads/s %.1f'%(d['value']/1e6), 'inflate %.1f'%d['stages']['inflate']['ms'], 'parse %.1f'%d['stages']['parse']['ms'], 'valid', d['config']['counters']['valid'], 'nnz', d['config']['counters']['nnz'])"
done; done
