#!/usr/bin/env python
"""Generates tests/golden/: tiny inputs + the outputs of the UNMODIFIED reference (oracle/_ref/fastF_ref, built by
oracle/Makefile from /root/reference/src) on them.  Run in the dev container (the reference sources do not travel)."""
import gzip
import json
import os
import shutil
import subprocess
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bamgen   # noqa: E402
import synth_binding   # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.path.join(ROOT, "oracle", "_ref", "fastF_ref")


def db_digest(path):
    """schema text + sha256 of every table's rows (in rowid order) of a bam2db sqlite file"""
    import hashlib
    import sqlite3
    c = sqlite3.connect(path)
    out = {"schema": [list(r) for r in c.execute("select name, sql from sqlite_master order by name")]}
    for (name,) in c.execute("select name from sqlite_master where type='table' order by name").fetchall():
        h = hashlib.sha256()
        n = 0
        for row in c.execute("select * from %s order by rowid" % name):
            h.update(repr(row).encode())
            n += 1
        out[name] = [n, h.hexdigest()]
    c.close()
    return out


def run_ref_bam2db(d, expect, rc, rd, seed, umicopies=False):
    out = os.path.join(d, expect)
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    db = os.path.join(d, "tmp.db")
    if os.path.exists(db):
        os.remove(db)
    r = subprocess.run([REF, "bam2db", "-b", "in.bam", "-f", "features.tsv.gz", "-a", "barcodes.tsv.gz", "-d", "tmp.db", "-c", str(rc), "-r", str(rd), "-o", expect, "-s", str(seed)] + (["-u"] if umicopies else []),
                       cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, check=True)
    json.dump(db_digest(db), open(os.path.join(out, "db_digest.json"), "w"), indent=1)
    os.remove(db)
    c = {}
    for ln in r.stdout.splitlines():
        if "total fastQ reads" in ln:
            c[0] = int(ln.rsplit(":", 1)[1])
        elif "sampled and valid" in ln:
            c[2] = int(ln.rsplit(":", 1)[1])
        elif "sampled fastQ reads" in ln:
            c[1] = int(ln.rsplit(":", 1)[1])
    return [c[0], c[1], c[2]]


def main():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    S = synth_binding.load()
    cases = []
    # 1. synthetic 10x-v3 shape, 4k reads
    d = os.path.join(GOLD, "synth4k")
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    paths, _ = S.write_bam_set(d, n_reads=4000, n_cells=60, n_genes=120, seed=11, p_umi_n=0.01, n_molecules=1500)
    os.rename(paths["bam"], os.path.join(d, "in.bam"))
    for rc, rd, seed in ((0.5, 0.5, 926), (1.0, 0.3, 926), (1.0, 1.0, 1), (0.2, 0.9, 77)):
        exp = "expect_c%s_r%s_s%d" % (rc, rd, seed)
        cnt = run_ref_bam2db(d, exp, rc, rd, seed, umicopies=True)
        cases.append({"name": "synth4k-c%s-r%s-s%d" % (rc, rd, seed), "kind": "bam2db", "dir": "synth4k", "expect": exp, "rate_cell": rc, "rate_depth": rd, "seed": seed, "counters": cnt, "umicopies": True})
    # 2. hand-crafted edge cases, every deflate block type, a multi-block header
    d = os.path.join(GOLD, "edge")
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    import random
    rng = random.Random(5)
    barcodes = sorted({"".join(rng.choice("ACGT") for _ in range(16)) + "-1" for _ in range(24)})
    genes = ["ENSG%011d" % (i + 1) for i in range(40)]
    with gzip.open(os.path.join(d, "barcodes.tsv.gz"), "wb") as f:
        f.write("".join(b + "\n" for b in barcodes).encode())
    with gzip.open(os.path.join(d, "features.tsv.gz"), "wb") as f:
        f.write("".join("%s\tGene%d\tGene Expression\n" % (g, i) for i, g in enumerate(genes)).encode())
    recs = bamgen.edge_case_bam(barcodes, genes)
    header = bamgen.bam_header(text=b"@HD\tVN:1.6\tSO:coordinate\n" + b"@CO\t" + b"x" * 70000 + b"\n", refs=[(b"chr%d" % i, 1000000 + i) for i in range(30)])
    chunks = bamgen.pack_records(header, recs, max_payload=9000)
    modes = [(6, zlib.Z_DEFAULT_STRATEGY), (0, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (9, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)]
    open(os.path.join(d, "in.bam"), "wb").write(bamgen.bgzf_file(chunks, modes))
    for rc, rd, seed in ((1.0, 1.0, 926), (0.5, 0.5, 926), (0.8, 0.6, 3)):
        exp = "expect_c%s_r%s_s%d" % (rc, rd, seed)
        cnt = run_ref_bam2db(d, exp, rc, rd, seed)
        cases.append({"name": "edge-c%s-r%s-s%d" % (rc, rd, seed), "kind": "bam2db", "dir": "edge", "expect": exp, "rate_cell": rc, "rate_depth": rd, "seed": seed, "counters": cnt})
    # 3. freq: synthetic + ragged text (N, short lines, missing final newline, truncated last record)
    d = os.path.join(GOLD, "freq")
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    fq, _ = S.write_fastq(d, n_reads=3000, n_cells=40, seed=4, p_umi_n=0.02)
    os.rename(fq, os.path.join(d, "synth.fastq.gz"))
    rng = random.Random(9)
    lines = []
    for i in range(400):
        r = rng.random()
        seq = "".join(rng.choice("ACGT") for _ in range(28)) if rng.random() < 0.5 else rng.choice(["ACGTACGTACGTACGTAAAACCCCGGGG", "TTTTTTTTTTTTTTTTTTTTTTTTTTTT"])
        if r < 0.05:
            seq = seq[:rng.randrange(0, 27)]                 # short read: the key swallows the newline
        elif r < 0.12:
            k = rng.randrange(28)
            seq = seq[:k] + "N" + seq[k + 1:]
        elif r < 0.14:
            seq = seq.lower()
        lines.append("@r%d\n%s\n+\n%s\n" % (i, seq, "F" * len(seq)))
    ragged = "".join(lines)
    variants = {"ragged.fastq.gz": ragged, "ragged_nonl.fastq.gz": ragged[:-1], "ragged_trunc.fastq.gz": ragged + "@last\nACGTACGTACGTACGTACGTACGTACGT\n"}
    for name, text in variants.items():
        raw = text.encode()
        blocks = [raw[i:i + 5000] for i in range(0, len(raw), 5000)]
        open(os.path.join(d, name), "wb").write(bamgen.bgzf_file(blocks))
    for name in ["synth.fastq.gz"] + list(variants):
        for l, u in ((16, 12), (16, 0), (5, 3)):
            out = os.path.join(d, "tmpout")
            shutil.rmtree(out, ignore_errors=True)
            os.makedirs(out)
            subprocess.run([REF, "freq", "-R", name, "-o", "tmpout", "-l", str(l), "-u", str(u)], cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            exp = "expect_%s_l%d_u%d.txt.gz" % (name.split(".")[0], l, u)
            with open(os.path.join(out, "whitelist.txt"), "rb") as fi, gzip.open(os.path.join(d, exp), "wb") as fo:
                fo.write(fi.read())
            shutil.rmtree(out)
            cases.append({"name": "freq-%s-l%d-u%d" % (name.split(".")[0], l, u), "kind": "freq", "dir": "freq", "input": name, "expect": exp, "l": l, "u": u})
    json.dump({"generated_by": "scripts/make_golden.py (unmodified reference compiled by oracle/Makefile)", "cases": cases}, open(os.path.join(GOLD, "manifest.json"), "w"), indent=1)
    print("wrote", len(cases), "cases;", subprocess.run(["du", "-sh", GOLD], stdout=subprocess.PIPE, text=True).stdout.strip())


if __name__ == "__main__":
    main()
