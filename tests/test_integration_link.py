"""INTEGRATION.md section B as a test: the reference's UNMODIFIED src/main.c + argparse.c (compiled by oracle/Makefile where they lie
under /root/reference) linked against this repository's bam2db() (fastf_b200/host).  `fastF_gpu bam2db ...` goes through the reference's
own option parsing and pre-checks (src/main.c:288-362) into our drop-in; outputs must equal the goldens recorded from the reference.
CPU: linked against the SIMT-emulator build of the kernels.  GPU (-m gpu): linked against libfastf_gpu.so."""
import gzip
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(ROOT, "tests"))
CASES = [c for c in json.load(open(os.path.join(GOLD, "manifest.json")))["cases"] if c["kind"] == "bam2db"]


def _run_case(exe, case, out):
    from dbdigest import db_digest
    d = os.path.join(GOLD, case["dir"])
    cmd = [exe, "bam2db", "-b", "in.bam", "-f", "features.tsv.gz", "-a", "barcodes.tsv.gz", "-d", os.path.join(out, "x.db"), "-c", str(case["rate_cell"]), "-r", str(case["rate_depth"]),
           "-o", out, "-s", str(case["seed"])] + (["-u"] if case.get("umicopies") else [])
    r = subprocess.run(cmd, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "In in.bam, total fastQ reads: %d" % case["counters"][0] in r.stdout
    for f in ["matrix.mtx.gz", "barcodes.tsv.gz", "features.tsv.gz"] + (["umi.tsv.gz"] if case.get("umicopies") else []):
        assert gzip.open(os.path.join(out, f), "rb").read() == gzip.open(os.path.join(d, case["expect"], f), "rb").read(), f
    assert db_digest(os.path.join(out, "x.db")) == json.load(open(os.path.join(d, case["expect"], "db_digest.json")))
    # the reference's own pre-check still guards the database (src/main.c:341-345)
    r = subprocess.run(cmd, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode != 0 and "already exists" in r.stderr


@pytest.mark.parametrize("name", ["synth4k-c0.5-r0.5-s926", "synth4k-c1.0-r0.3-s926", "edge-c0.8-r0.6-s3"])
def test_reference_main_with_our_bam2db_on_the_emulator(name, tmp_path):
    from fastf_b200 import build
    if not os.path.isdir("/root/reference/src") and not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "fastF_gpu_emu")):
        pytest.skip("reference sources absent and no prebuilt fastF_gpu_emu")
    exe = build.build_ref_main_emu()
    assert exe, "oracle/_ref/fastF_gpu_emu was not built"
    case = [c for c in CASES if c["name"] == name]
    if not case:
        pytest.skip("golden case %s not in the manifest" % name)
    _run_case(exe, case[0], str(tmp_path))


@pytest.mark.gpu
@pytest.mark.parametrize("name", [c["name"] for c in CASES])
def test_reference_main_with_our_bam2db_on_the_gpu(name, tmp_path):
    exe = os.path.join(ROOT, "oracle", "_ref", "fastF_gpu")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/fastF_gpu not built (reference sources were absent at build time)")
    _run_case(exe, [c for c in CASES if c["name"] == name][0], str(tmp_path))
