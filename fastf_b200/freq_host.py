"""Host side of `fastF freq`, mirroring the reference's

    node *cell_counts(gzFile R1_file, size_t len_cellbarcode, size_t len_umi)       (reference src/count.h:6)
    print_tree(node*, FILE*)  ->  <out>/whitelist.txt, one "key,count\\n" per key    (src/main.c:70-89, src/filter.c:139-148)

The histogram is computed on the GPU (libfastf_gpu.so: fastf_freq_gpu); this module merges the handful of reads whose
key is not pure ACGT (exported raw by the device) and writes the file in the reference's BST pre-order."""
import ctypes as C
import os
import sys

import numpy as np

from . import _lib


def _decode_keys(keys, klen):
    """2-bit keys (first base most significant) -> (n, klen) uint8 array of ACGT bytes"""
    k = np.asarray(keys, dtype=np.uint64)
    shifts = np.arange(klen - 1, -1, -1, dtype=np.uint64) * np.uint64(2)
    codes = ((k[:, None] >> shifts[None, :]) & np.uint64(3)).astype(np.uint8)
    return np.frombuffer(b"ACGT", dtype=np.uint8)[codes]


class Histogram:
    """Distinct keys in ascending strcmp order with their counts and first-occurrence read ordinals."""

    def __init__(self, keys_bytes, count, first, n_reads):
        self.keys = keys_bytes      # list/array of bytes objects (ascending)
        self.count = count
        self.first = first
        self.n_reads = n_reads


def cell_counts(R1_file, len_cellbarcode, len_umi, device=0, ctx=None, inflate_lanes=0, stats_out=None):
    """Reference-shaped entry point: returns the histogram (the BST's content) of the first l+u bases of every read."""
    own = ctx is None
    if own:
        ctx = _lib.Context(device)
    try:
        lib = ctx.lib
        data = np.fromfile(R1_file, dtype=np.uint8) if isinstance(R1_file, (str, os.PathLike)) else np.frombuffer(R1_file, dtype=np.uint8)
        klen = int(len_cellbarcode) + int(len_umi)
        res = _lib.FreqResult()
        ctx.check(lib.fastf_freq_gpu(ctx.h, C.c_void_p(data.ctypes.data), data.size, klen, inflate_lanes, C.byref(res)), "freq_gpu")
        try:
            n = int(res.n_keys)
            key = np.ctypeslib.as_array(res.key, (n,)).copy() if n else np.zeros(0, np.uint64)
            count = np.ctypeslib.as_array(res.count, (n,)).astype(np.int64) if n else np.zeros(0, np.int64)
            first = np.ctypeslib.as_array(res.first, (n,)).astype(np.int64) if n else np.zeros(0, np.int64)
            ne = int(res.n_exceptions)
            stride = int(res.exc_stride)
            exc_ord = np.ctypeslib.as_array(res.exc_ordinal, (ne,)).copy() if ne else np.zeros(0, np.uint32)
            exc_raw = np.ctypeslib.as_array(res.exc_bytes, (ne * stride,)).copy().reshape(ne, stride) if ne else np.zeros((0, stride), np.uint8)
            n_reads = int(res.n_reads)
            n_dev_records = n + 0
            if stats_out is not None:
                stats_out.update({f: getattr(res, f) for f, t in _lib.FreqResult._fields_ if t in (C.c_uint64, C.c_uint32, C.c_float, C.c_uint8)})
        finally:
            lib.fastf_freq_result_free(C.byref(res))
        dec = _decode_keys(key, klen)
        keys_b = [bytes(r) for r in dec] if n < 200000 else dec.view("S%d" % klen).ravel().tolist()
        # exceptional reads: key = bytes of the sequence line up to klen, cut after the first '\n' (strncpy stops at the NUL that
        # follows it in gzgets' buffer, src/filter.c:270) or at a NUL byte
        extra = {}
        for o, raw in zip(exc_ord.tolist(), exc_raw):
            b = bytes(raw[:klen])
            z = b.find(b"\0")
            if z >= 0:
                b = b[:z]
            nl = b.find(b"\n")
            if nl >= 0:
                b = b[:nl + 1]
            c, f = extra.get(b, (0, o))
            extra[b] = (c + 1, min(f, o))
        # a trailing record whose id line exists but whose sequence line does not: empty key
        n_counted = int(count.sum()) + len(exc_ord)
        if n_reads > n_counted:
            c, f = extra.get(b"", (0, n_counted))
            extra[b""] = (c + (n_reads - n_counted), min(f, n_counted))
        if extra:
            allk = keys_b + list(extra.keys())
            allc = np.concatenate([count, np.array([v[0] for v in extra.values()], dtype=np.int64)])
            allf = np.concatenate([first, np.array([v[1] for v in extra.values()], dtype=np.int64)])
            order = sorted(range(len(allk)), key=lambda i: allk[i])
            keys_b = [allk[i] for i in order]
            count = allc[order]
            first = allf[order]
        return Histogram(keys_b, count, first, n_reads)
    finally:
        if own:
            ctx.close()


def print_tree(hist, fp):
    """Writes "key,count\\n" lines in the pre-order of the reference's BST (src/filter.c:139-148)."""
    n = len(hist.keys)
    if n == 0:
        return
    lib = _lib.load()
    first = np.ascontiguousarray(hist.first, dtype=np.uint32)
    order = np.zeros(n, dtype=np.uint64)
    if lib.fastf_cartesian_preorder(first.ctypes.data_as(_lib.c_u32p), n, order.ctypes.data_as(_lib.c_u64p)) != 0:
        raise _lib.FastfError("cartesian_preorder failed")
    cnt = hist.count
    keys = hist.keys
    fp.write(b"".join(keys[i] + b"," + str(int(cnt[i])).encode() + b"\n" for i in order.tolist()))


def freq(R1, out_dir, len_cellbarcode=16, len_umi=10, device=0, ctx=None):
    """`fastF freq -R R1 -o out_dir -l L -u U` (reference src/main.c:30-92): writes <out_dir>/whitelist.txt; returns 0 / 1."""
    try:
        if not os.path.exists(R1):
            sys.stderr.write("Can't open R1 file %s\n" % R1)
            return 1
        hist = cell_counts(R1, len_cellbarcode, len_umi, device=device, ctx=ctx)
        with open(os.path.join(out_dir, "whitelist.txt"), "wb") as f:
            print_tree(hist, f)
        return 0
    except (_lib.FastfError, OSError, ValueError) as e:
        sys.stderr.write("freq: %s\n" % e)
        return 1
