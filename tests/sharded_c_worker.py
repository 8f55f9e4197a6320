"""TEST INFRASTRUCTURE: the one-process multi-GPU driver of the library (fastf_bam2db_run_sharded) against the single-job result and
the CPU oracle on the same seeded input.  Runs on whatever FASTF_GPU_LIB names (the emulator build with FASTF_EMU_DEVICES "devices" on
CPU; libfastf_gpu.so on a multi-GPU box).  usage: sharded_c_worker.py <n_devices> <n_reads> <rate_cell> <rate_depth> <seed>"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from fastf_b200 import _lib, bam2db_host as B   # noqa: E402
import oracle_binding   # noqa: E402
import synth_binding   # noqa: E402

G, n_reads, rc, rd, seed = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5])
O, S = oracle_binding.load(), synth_binding.load()
lib = _lib.load()
with tempfile.TemporaryDirectory() as d:
    paths, _ = S.write_bam_set(d, n_reads=n_reads, n_cells=max(20, n_reads // 300), n_genes=max(30, n_reads // 150), seed=seed, p_umi_n=0.01, n_molecules=max(100, n_reads // 3))
    want = O.bam2db(paths["bam"], paths["barcodes"], paths["features"], rc, rd, seed)
    inputs = B.Bam2dbInputs(lib, paths["barcodes"], paths["features"], rc, seed)
    bam = np.fromfile(paths["bam"], dtype=np.uint8)
    with _lib.Context(0) as ctx:
        one_stats, one = B.run_device(ctx, bam, inputs, rd, seed, want_rows=True)
    for g in (sorted({1, G}) if n_reads <= 3000 or n_reads >= 100000 else [G]):
        st, out = B.run_sharded(lib, bam, inputs, rd, seed, g, want_rows=True)
        for k in ("total", "cb_valid", "sampled", "valid", "nnz"):
            assert st[k] == want[k] == one_stats[k], (g, k, st[k], want[k], one_stats[k])
        for k in ("m_gene", "m_cell", "m_count"):
            assert np.array_equal(out[k], want[k]) and np.array_equal(out[k], one[k]), (g, k)
        assert np.array_equal(out["row_keys"], one["row_keys"]), (g, "rows")
        assert g == 1 or st["exchanged_keys"] > 0
        # without rows the result must not change
        st2, out2 = B.run_sharded(lib, bam, inputs, rd, seed, g, want_rows=False)
        assert st2["nnz"] == st["nnz"] and np.array_equal(out2["m_count"], out["m_count"]) and out2["row_keys"].size == 0
    # a shard that fails (garbage instead of BGZF) fails the call with a message, it does not hang or crash
    bad = bam.copy()
    bad[bad.size // 2: bad.size // 2 + 4096] = 7
    try:
        B.run_sharded(lib, bad, inputs, rd, seed, G, want_rows=False)
        raise SystemExit("corrupt input was accepted")
    except _lib.FastfError as e:
        assert "run_sharded" in str(e)
print("OK sharded C driver: %d devices, %d reads, nnz %d, exchanged %d keys" % (G, want["total"], want["nnz"], st["exchanged_keys"]))
