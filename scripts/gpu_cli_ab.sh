#!/bin/bash
# A/B of the whole CLI run between two builds of the tree (the current one and a worktree of an older commit) on the same box:
# is a change in `fastF bam2db` wall time ours or the box's (CUDA initialisation without a persistence daemon)?
# usage: gpu_cli_ab.sh <old-worktree-dir> [reads=5000000]
OLD=$1; READS=${2:-5000000}
D=/dev/shm/fastf_cliab_$$; mkdir -p $D gpurun_out
nvidia-smi -q | grep -i "persistence mode" | head -1
fastf_b200/_build/fastf_synth bam --out $D --reads $READS --cells 10000 --genes 36000 --seed 4242 2>&1 | tail -1
ls $D | head
for round in 1 2; do
  for tag in new old; do
    bin=fastf_b200/_build/fastF; [ $tag = old ] && bin=$OLD/fastf_b200/_build/fastF
    rm -rf $D/out_$tag $D/x_$tag.db; mkdir -p $D/out_$tag
    t0=$(date +%s%N)
    FASTF_HOST_TIMING=1 $bin bam2db -b $D/synth.bam -f $D/features.tsv.gz -a $D/barcodes.tsv.gz -d $D/x_$tag.db -c 1.0 -r 0.3 -o $D/out_$tag -s 926 > $D/log_$tag.txt 2>&1
    t1=$(date +%s%N)
    echo "$tag round $round: $(( (t1 - t0) / 1000000 )) ms wall; $(grep 'device job' $D/log_$tag.txt)"
  done
done
rm -rf $D
