#!/usr/bin/env python
"""Generates tests/golden/tags/: a hand-crafted BAM for `crb` / `extract` plus the outputs of the UNMODIFIED reference
(oracle/_ref/fastF_ref) on it and on tests/golden/synth4k/in.bam.  Run in the dev container (the reference sources do not travel).
The BAM avoids the inputs on which the reference dereferences NULL (CB without CR, string extraction of a non-string tag)."""
import gzip
import json
import os
import random
import shutil
import subprocess
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bamgen as G   # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.path.join(ROOT, "oracle", "_ref", "fastF_ref")


def tags_bam(seed=13):
    rng = random.Random(seed)
    cbs = ["".join(rng.choice("ACGT") for _ in range(16)) + "-1" for _ in range(30)] + ["NNNNACGTACGTACGT-1", "short", ""]
    recs = []
    n = 0

    def add(aux, **kw):
        nonlocal n
        recs.append(G.record("q%06d" % n, aux, rng=rng, pos=n, **kw))
        n += 1

    def cr_of(cb):
        base = cb[:16]
        r = rng.random()
        if r < 0.7 or not base:
            return base
        k = rng.randrange(len(base))
        return base[:k] + rng.choice("ACGTN") + base[k + 1:]          # sequencing error in the raw barcode

    for i in range(900):
        cb = rng.choice(cbs)
        aux = [G.aux_int("NH", rng.choice("cCsSiI"), rng.choice([1, 1, 1, 2, 3, 10])), G.aux_int("xf", "C", rng.choice([25, 25, 17, 0])),
               G.aux_int("AS", "s", rng.choice([-3, 89, 90, 255, 999])), G.aux_Z("GX", "ENSG%011d" % rng.randrange(1, 50)), G.aux_Z("RG", "sample:0:1:HXXX:%d" % rng.randrange(3))]
        r = rng.random()
        if r < 0.06:
            pass                                                     # no CB, no CR
        elif r < 0.1:
            aux.append(G.aux_Z("CR", cr_of(cb)))                     # CR without CB: ignored by crb
        else:
            aux += [G.aux_Z("CR", cr_of(cb)), G.aux_Z("CY", "F" * 16), G.aux_Z("CB", cb)]
        if rng.random() < 0.1:
            aux.insert(0, G.aux_Z("CB", rng.choice(cbs)))            # duplicate tag: the first one counts
            aux.insert(0, G.aux_Z("CR", "ACGTACGTACGTACGT"))
        if rng.random() < 0.05:
            aux.insert(0, G.aux_B("ZB", "S", [1, 2, 3]))
        add(aux)
    add([G.aux_Z("CR", "A" * 16), G.aux_Z("CB", "A" * 16 + "-1"), G.aux_int("NH", "I", 4294967295), G.aux_int("xf", "i", -7), G.aux_Z("GX", "x" * 200)])
    add([G.aux_Z("ZL", "y" * 14000), G.aux_Z("CR", "C" * 16), G.aux_Z("CB", "C" * 16 + "-1"), G.aux_int("NH", "C", 1), G.aux_Z("GX", "behind-a-record-larger-than-the-window")])
    add([G.aux_H("CB", "1AE3"), G.aux_H("CR", "00FF"), G.aux_A("NH", "x"), G.aux_f("xf", 2.5)])     # H strings are strings; A / f read as integer 0
    add([])
    header = G.bam_header(text=b"@HD\tVN:1.6\n@CO\tcrb/extract fixtures\n", refs=[(b"chr1", 1000)])
    chunks = G.pack_records(header, recs, max_payload=20000)
    modes = [(6, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (0, zlib.Z_DEFAULT_STRATEGY)]
    return G.bgzf_file(chunks, modes)


def main():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    d = os.path.join(GOLD, "tags")
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    open(os.path.join(d, "tags.bam"), "wb").write(tags_bam())
    cases = []
    for bam, stem in (("tags.bam", "tags"), ("../synth4k/in.bam", "synth4k")):
        r = subprocess.run([REF, "crb", "-b", bam, "-o", "tmp.gz"], cwd=d, check=True, stdout=subprocess.PIPE, text=True)
        reads = int([ln for ln in r.stdout.splitlines() if ln.startswith("Processed all")][0].split()[2])
        exp = "expect_%s_crb.txt.gz" % stem
        with gzip.open(os.path.join(d, "tmp.gz"), "rb") as fi, gzip.open(os.path.join(d, exp), "wb") as fo:
            fo.write(fi.read())
        os.remove(os.path.join(d, "tmp.gz"))
        cases.append({"name": "crb-%s" % stem, "kind": "crb", "input": bam, "expect": exp, "reads": reads})
        for tag, typ in (("GX", 0), ("CB", 0), ("RG", 0), ("xf", 1), ("NH", 1), ("AS", 1), ("CB", 1), ("ZZ", 0)):
            if stem == "synth4k" and tag in ("RG", "AS", "ZZ"):
                continue
            r = subprocess.run([REF, "extract", "-b", bam, "-t", tag, "-T", str(typ)], cwd=d, check=True, stdout=subprocess.PIPE, text=True)
            tot = int([ln for ln in r.stdout.splitlines() if ln.startswith("Processed all")][0].split()[2])
            val = int([ln for ln in r.stdout.splitlines() if ln.startswith("Valid reads")][0].split()[2])
            exp = "expect_%s_extract_%s_%d.csv.gz" % (stem, tag, typ)
            with open(os.path.join(d, "tag_summary.csv"), "rb") as fi, gzip.open(os.path.join(d, exp), "wb") as fo:
                fo.write(fi.read())
            os.remove(os.path.join(d, "tag_summary.csv"))
            cases.append({"name": "extract-%s-%s-%d" % (stem, tag, typ), "kind": "extract", "input": bam, "tag": tag, "type": typ, "expect": exp, "total": tot, "valid": val})
    json.dump({"generated_by": "scripts/make_golden_tags.py (unmodified reference compiled by oracle/Makefile)", "cases": cases}, open(os.path.join(d, "manifest.json"), "w"), indent=1)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
