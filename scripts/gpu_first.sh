#!/bin/bash
# first contact with the B200: config-1-shaped bam2db and a 4M-read freq against the compiled reference
R=$(pwd)
T=/tmp/g1
mkdir -p $T/out $T/ref $T/fout $T/fref gpurun_out
{
nproc; grep -m1 "model name" /proc/cpuinfo; free -g | head -2; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
ls -la fastf_b200/_build oracle/_ref oracle/_build
time fastf_b200/_build/fastf_synth bam --out $T --reads 1000000 --cells 1000 --genes 2000 --umi-n 0.001
time fastf_b200/_build/fastf_synth fastq --out $T --reads 4000000 --cells 20000 --umi-n 0.001
( time oracle/_ref/fastF_ref bam2db -b $T/synth.bam -f $T/features.tsv.gz -a $T/barcodes.tsv.gz -d $T/ref.db -c 0.5 -r 0.5 -o $T/ref -s 926 ) 2>&1 | tail -12
( time oracle/_ref/fastF_ref freq -R $T/R1.fastq.gz -o $T/fref -l 16 -u 12 ) 2>&1 | tail -5
python - <<PY
import time, fastf_b200, numpy as np, ctypes as C
from fastf_b200 import _lib, bam2db_host as B
ctx = _lib.Context(0)
for rep in range(2):
    t=time.time()
    import os
    if os.path.exists('$T/our.db'): os.remove('$T/our.db')
    rc = fastf_b200.bam2db('$T/synth.bam','$T/our.db','$T/out','$T/barcodes.tsv.gz','$T/features.tsv.gz',0.5,0.5,926, ctx=ctx)
    print('bam2db rc',rc, 'wall', time.time()-t, flush=True)
inputs = B.Bam2dbInputs(ctx.lib, '$T/barcodes.tsv.gz','$T/features.tsv.gz',0.5,926)
bam = np.fromfile('$T/synth.bam', dtype=np.uint8)
for lanes in (32,16,8):
    for rep in range(3):
        t=time.time()
        stats,out = B.run_device(ctx, bam, inputs, 0.5, 926, want_rows=False, inflate_lanes=lanes)
        w=time.time()-t
    print('lanes',lanes,'wall',w, {k:(round(v,3) if isinstance(v,float) else v) for k,v in stats.items()}, flush=True)
for rep in range(2):
    t=time.time()
    rc = fastf_b200.freq('$T/R1.fastq.gz','$T/fout',16,12, ctx=ctx)
    print('freq rc',rc,'wall',time.time()-t, flush=True)
st={}
from fastf_b200 import freq_host as F
t=time.time(); h=F.cell_counts('$T/R1.fastq.gz',16,12,ctx=ctx,stats_out=st); print('cell_counts wall',time.time()-t, st)
PY
for f in matrix.mtx barcodes.tsv features.tsv; do cmp <(zcat $T/ref/$f.gz) <(zcat $T/out/$f.gz) && echo "$f identical"; done
cmp $T/fref/whitelist.txt $T/fout/whitelist.txt && echo "whitelist identical"
} > gpurun_out/first.log 2>&1
tail -40 gpurun_out/first.log
