"""world_size-2 CPU test (gloo) of the multi-GPU path: the sharded job must reproduce the reference's recorded output byte for byte."""
import gzip
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_shard_ranges():
    from fastf_b200.sharded import shard_ranges
    assert shard_ranges(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert shard_ranges(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert shard_ranges(0, 2) == [(0, 0), (0, 0)]


@pytest.mark.parametrize("name,world", [("edge-c0.5-r0.5-s926", 2), ("synth4k-c0.5-r0.5-s926", 3)])
def test_sharded_equals_reference(name, world, tmp_path):
    from fastf_b200 import build
    from dbdigest import db_digest
    emu = build.build_emu()
    case = [c for c in json.load(open(os.path.join(GOLD, "manifest.json")))["cases"] if c["name"] == name][0]
    d = os.path.join(GOLD, case["dir"])
    # FASTF_MT_JUMP_MIN=0: later ranks reach their first draw by GF(2) jump-ahead even on these tiny inputs
    env = dict(os.environ, FASTF_GPU_LIB=emu, OMP_NUM_THREADS="1", FASTF_MT_JUMP_MIN="0")
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "sharded_worker.py"), d, str(tmp_path), str(case["rate_cell"]), str(case["rate_depth"]), str(case["seed"])]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]
    for f in ("matrix.mtx.gz", "barcodes.tsv.gz", "features.tsv.gz"):
        assert gzip.open(os.path.join(tmp_path, f), "rb").read() == gzip.open(os.path.join(d, case["expect"], f), "rb").read(), f
    want = json.load(open(os.path.join(d, case["expect"], "db_digest.json")))
    got = db_digest(os.path.join(tmp_path, "x.db"))
    for t in ("cell", "feature", "umi", "mtx"):
        assert got[t] == want[t], t
