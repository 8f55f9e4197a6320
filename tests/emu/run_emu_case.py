"""TEST INFRASTRUCTURE: runs the Python host against the SIMT-emulator build of the kernels (FASTF_GPU_LIB must point at
tests/emu/_build/libfastf_emu.so) on one golden case and compares with the reference's recorded output."""
import gzip
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
assert "libfastf_emu" in os.environ.get("FASTF_GPU_LIB", ""), "refusing to run: this script is for the emulator build only"
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fastf_b200   # noqa: E402
from fastf_b200 import bam2db_host   # noqa: E402
from dbdigest import db_digest   # noqa: E402
import subprocess   # noqa: E402

USE_CLI = os.environ.get("FASTF_EMU_CLI")   # path of the C host linked against the emulator build

GOLD = os.path.join(ROOT, "tests", "golden")
cases = {c["name"]: c for c in json.load(open(os.path.join(GOLD, "manifest.json")))["cases"]}
bad = 0
for name in sys.argv[1:]:
    c = cases[name]
    d = os.path.join(GOLD, c["dir"])
    with tempfile.TemporaryDirectory() as t:
        if c["kind"] == "bam2db":
            cwd = os.getcwd()
            os.chdir(d)   # the header records the BAM path as passed: the golden run used "in.bam"
            try:
                if USE_CLI:
                    cmd = [USE_CLI, "bam2db", "-b", "in.bam", "-f", "features.tsv.gz", "-a", "barcodes.tsv.gz", "-d", os.path.join(t, "x.db"), "-c", str(c["rate_cell"]), "-r", str(c["rate_depth"]),
                           "-o", t, "-s", str(c["seed"])] + (["-u"] if c.get("umicopies") else [])
                    rc = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE).returncode
                else:
                    bam2db_host._umi_copies_flag = 1 if c.get("umicopies") else 0
                    rc = fastf_b200.bam2db("in.bam", os.path.join(t, "x.db"), t, "barcodes.tsv.gz", "features.tsv.gz", c["rate_cell"], c["rate_depth"], c["seed"])
            finally:
                os.chdir(cwd)
            ok = rc == 0
            files = ["matrix.mtx.gz", "barcodes.tsv.gz", "features.tsv.gz"] + (["umi.tsv.gz"] if c.get("umicopies") else [])
            for f in files:
                ok = ok and gzip.open(os.path.join(t, f), "rb").read() == gzip.open(os.path.join(d, c["expect"], f), "rb").read()
            ok = ok and db_digest(os.path.join(t, "x.db")) == json.load(open(os.path.join(d, c["expect"], "db_digest.json")))
        else:
            if USE_CLI:
                rc = subprocess.run([USE_CLI, "freq", "-R", os.path.join(d, c["input"]), "-o", t, "-l", str(c["l"]), "-u", str(c["u"])], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE).returncode
            else:
                rc = fastf_b200.freq(os.path.join(d, c["input"]), t, c["l"], c["u"])
            ok = rc == 0 and open(os.path.join(t, "whitelist.txt"), "rb").read() == gzip.open(os.path.join(d, c["expect"]), "rb").read()
    print(("PASS " if ok else "FAIL ") + name, flush=True)
    bad += not ok
sys.exit(1 if bad else 0)
