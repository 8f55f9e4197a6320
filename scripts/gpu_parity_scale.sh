#!/bin/bash
# Parity where the benchmark runs: `fastF bam2db` (this repository, default chunking, file fed in 256 MiB pieces) against the UNMODIFIED
# reference CLI (oracle/_ref/fastF_ref) on N distinct synthetic reads of the BASELINE configs[2] shape (10k cells, 36k genes, -c 1.0 -r 0.3 -s 926).
# Compares the decompressed matrix.mtx / barcodes / features and the sha256 of every row of every sqlite table.
# usage: gpu_parity_scale.sh [reads=64000000] [gpus=1]      log -> gpurun_out/parity_scale_<reads>_g<gpus>.log
READS=${1:-64000000}; GPUS=${2:-1}
mkdir -p gpurun_out
LOG=gpurun_out/parity_scale_${READS}_g${GPUS}.log
D=/dev/shm/fastf_scale_$$; rm -rf $D; mkdir -p $D/ref $D/ours
{
echo "== parity at scale: $READS distinct reads, $GPUS GPU(s), $(nproc) host cores, $(date -u +%FT%TZ)"
T0=$(date +%s%N); fastf_b200/_build/fastf_synth bam --out $D --reads $READS --cells 10000 --genes 36000 --seed 4242 2>&1 | tail -2; echo "generate: $(( ($(date +%s%N) - T0) / 1000000 )) ms"
ls -l $D/synth.bam | awk '{print "BAM bytes:", $5}'
T0=$(date +%s%N); oracle/_ref/fastF_ref bam2db -b $D/synth.bam -f $D/features.tsv.gz -a $D/barcodes.tsv.gz -d $D/ref/x.db -c 1.0 -r 0.3 -o $D/ref -s 926 2>&1 | grep -v "^Opened\|generated" | tail -6; echo "reference CLI: $(( ($(date +%s%N) - T0) / 1000000 )) ms wall"
T0=$(date +%s%N); FASTF_HOST_TIMING=1 FASTF_GPUS=$GPUS fastf_b200/_build/fastF bam2db -b $D/synth.bam -f $D/features.tsv.gz -a $D/barcodes.tsv.gz -d $D/ours/x.db -c 1.0 -r 0.3 -o $D/ours -s 926 2>&1 | grep -v "^Opened\|generated" | tail -12; echo "fastF (B200) CLI: $(( ($(date +%s%N) - T0) / 1000000 )) ms wall"
python - $D <<'PY'
import gzip, hashlib, sqlite3, sys, time
sys.path.insert(0, "tests")
from dbdigest import db_digest
d = sys.argv[1]
ok = True
for f in ("matrix.mtx.gz", "barcodes.tsv.gz", "features.tsv.gz"):
    a, b = gzip.open(f"{d}/ref/{f}", "rb").read(), gzip.open(f"{d}/ours/{f}", "rb").read()
    print(f"{f}: {len(a)} bytes decompressed, sha256 {hashlib.sha256(a).hexdigest()[:16]} vs {hashlib.sha256(b).hexdigest()[:16]} -> {'IDENTICAL' if a == b else 'DIFFERENT'}")
    ok &= a == b
t = time.time()
A, B = db_digest(f"{d}/ref/x.db"), db_digest(f"{d}/ours/x.db")
for t_ in ("cell", "feature", "umi", "mtx"):
    print(f"table {t_}: {A[t_][0]} rows, sha256 {A[t_][1][:16]} vs {B[t_][0]} rows, {B[t_][1][:16]} -> {'IDENTICAL' if A[t_] == B[t_] else 'DIFFERENT'}")
    ok &= A[t_] == B[t_]
print("integrity_check (ours):", sqlite3.connect(f"{d}/ours/x.db").execute("PRAGMA integrity_check").fetchall())
print("PARITY", "GREEN" if ok else "RED", "(digests took %.0f s)" % (time.time() - t))
sys.exit(0 if ok else 1)
PY
echo "exit=$?"
} > $LOG 2>&1
rm -rf $D
cat $LOG
