#!/bin/bash
# dev helper: run the Python host against the SIMT-emulator build of the C-ABI and compare with the unmodified reference
set -e
R=/root/repo
T=${1:-/tmp/t2}
READS=${2:-3000}
rm -rf $T && mkdir -p $T/out $T/ref $T/fout $T/fref
python -c "from fastf_b200 import build; build.build_emu(); build.build_synth()"
$R/fastf_b200/_build/fastf_synth bam --out $T --reads $READS --cells 50 --genes 100 --umi-n 0.01 --molecules $((READS/2))
$R/fastf_b200/_build/fastf_synth fastq --out $T --reads $READS --cells 40 --umi-n 0.01
$R/oracle/_ref/fastF_ref bam2db -b $T/synth.bam -f $T/features.tsv.gz -a $T/barcodes.tsv.gz -d $T/ref.db -c 0.5 -r 0.5 -o $T/ref -s 926 > $T/ref.log 2>&1
$R/oracle/_ref/fastF_ref freq -R $T/R1.fastq.gz -o $T/fref -l 16 -u 12 > $T/fref.log 2>&1
export FASTF_GPU_LIB=$R/tests/emu/_build/libfastf_emu.so
time python - <<PY
import fastf_b200, time
t=time.time()
rc = fastf_b200.bam2db('$T/synth.bam','$T/our.db','$T/out','$T/barcodes.tsv.gz','$T/features.tsv.gz',0.5,0.5,926)
print('bam2db rc',rc, time.time()-t)
t=time.time()
rc = fastf_b200.freq('$T/R1.fastq.gz','$T/fout',16,12)
print('freq rc',rc, time.time()-t)
PY
for f in matrix.mtx barcodes.tsv features.tsv; do cmp <(zcat $T/ref/$f.gz) <(zcat $T/out/$f.gz) && echo "$f identical"; done
cmp $T/fref/whitelist.txt $T/fout/whitelist.txt && echo "whitelist identical"
