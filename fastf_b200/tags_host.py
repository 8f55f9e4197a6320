"""Host-side mirror of the reference's `crb` and `extract` operators (reference src/extract.c, src/main.c:231-286,364-402):
the per-record work -- inflate, record framing, aux walk, grouping equal values and finding each value's first occurrence -- runs in
fastf_taghist_gpu (CUDA); the host only orders the distinct values the way the reference's unbalanced BSTs print them (pre-order =
Cartesian tree of the first occurrences over the bytewise-sorted values, fastf_cartesian_preorder) and writes the files.
There is no CPU path: without the CUDA library / a device these functions raise."""
import ctypes as C
import gzip
import os

import numpy as np

from . import _lib

TAG_STRING, TAG_INT = 0, 1


def taghist(ctx, bam_bytes, tag_a, mode=TAG_STRING, tag_b=None, inflate_lanes=0, retry_straddle=False):
    """-> (stats, groups); groups = list of (value_a: bytes, value_b: bytes | None, count, first) in no particular order"""
    buf = np.frombuffer(bam_bytes, dtype=np.uint8) if not isinstance(bam_bytes, np.ndarray) else bam_bytes
    res = _lib.TaghistResult()
    rc = ctx.lib.fastf_taghist_gpu(ctx.h, C.c_void_p(buf.ctypes.data), buf.size, tag_a.encode(), mode, tag_b.encode() if tag_b else None, inflate_lanes, C.byref(res))
    if rc and retry_straddle and not (inflate_lanes & 0x400) and b"record-straddles-bgzf-block" in ctx.lib.fastf_last_error(ctx.h):
        # not an htslib-written file (records cross BGZF blocks): once more with FASTF_BAM_STRADDLE
        rc = ctx.lib.fastf_taghist_gpu(ctx.h, C.c_void_p(buf.ctypes.data), buf.size, tag_a.encode(), mode, tag_b.encode() if tag_b else None, inflate_lanes | 0x400, C.byref(res))
    ctx.check(rc, "taghist")
    try:
        n = int(res.n_groups)
        first = np.ctypeslib.as_array(res.first, (max(n, 1),))[:n].copy()
        count = np.ctypeslib.as_array(res.count, (max(n, 1),))[:n].copy()
        groups = []
        if mode == TAG_INT:
            iv = np.ctypeslib.as_array(res.ivalue, (max(n, 1),))[:n]
            groups = [(b"%d" % int(iv[g]), None, int(count[g]), int(first[g])) for g in range(n)]
        else:
            blob = C.string_at(res.strings, int(res.strings_bytes))
            a_off = np.ctypeslib.as_array(res.a_off, (max(n, 1),))[:n]
            a_len = np.ctypeslib.as_array(res.a_len, (max(n, 1),))[:n]
            b_len = np.ctypeslib.as_array(res.b_len, (max(n, 1),))[:n]
            for g in range(n):
                o, la, lb = int(a_off[g]), int(a_len[g]), int(b_len[g])
                groups.append((blob[o:o + la], blob[o + la:o + la + lb] if tag_b else None, int(count[g]), int(first[g])))
        stats = {f: getattr(res, f) for f, t in _lib.TaghistResult._fields_ if f not in ("first", "count", "ivalue", "a_off", "a_len", "b_len", "strings")}
        return stats, groups
    finally:
        ctx.lib.fastf_taghist_result_free(C.byref(res))


def _preorder(ctx, firsts):
    """positions (into the bytewise-sorted value list) in the order print_tree visits them (reference src/filter.c:139-148)"""
    n = len(firsts)
    t = np.ascontiguousarray(firsts, dtype=np.uint32)
    order = np.zeros(max(n, 1), dtype=np.uint64)
    if n:
        ctx.check(ctx.lib.fastf_cartesian_preorder(t.ctypes.data_as(_lib.c_u32p), n, order.ctypes.data_as(_lib.c_u64p)), "cartesian_preorder")
    return [int(x) for x in order[:n]]


def extract_bam(ctx, bam_file, tag, type_, out_dir="."):
    """reference extract_bam(bam_file, tag, type) (src/extract.c:135-216): writes <out_dir>/tag_summary.csv (the reference writes it into
    the working directory) and returns (total_count, valid_count) as the reference prints them -- total_count is doubled (:162,164)."""
    stats, groups = taghist(ctx, np.fromfile(bam_file, dtype=np.uint8), tag, TAG_INT if type_ else TAG_STRING, retry_straddle=True)
    groups.sort(key=lambda g: g[0])
    with open(os.path.join(out_dir, "tag_summary.csv"), "wb") as f:
        for i in _preorder(ctx, [g[3] for g in groups]):
            f.write(groups[i][0] + b",%d\n" % groups[i][2])
    return 2 * int(stats["n_records"]), int(stats["n_hits"])


def crb(ctx, bam_file, path_out):
    """reference cmd_crb: read_bam + print_CB_node (src/extract.c:47-133, src/main.c:231-286): one gz line "CB;CR,count;CR,count;...\\n"
    per cell barcode, CB nodes and the CR nodes under each in BST pre-order.  Returns read_count."""
    stats, groups = taghist(ctx, np.fromfile(bam_file, dtype=np.uint8), "CB", TAG_STRING, "CR", retry_straddle=True)
    by_cb = {}
    for cb, cr, count, first in groups:
        by_cb.setdefault(cb, []).append((cr, count, first))
    cbs = sorted(by_cb)
    with gzip.open(path_out, "wb") as f:
        for i in _preorder(ctx, [min(x[2] for x in by_cb[cb]) for cb in cbs]):
            crs = sorted(by_cb[cbs[i]])
            f.write(cbs[i] + b";" + b"".join(crs[j][0] + b",%d;" % crs[j][1] for j in _preorder(ctx, [x[2] for x in crs])) + b"\n")
    return int(stats["n_records"])
