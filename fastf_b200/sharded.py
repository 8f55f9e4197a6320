"""bam2db over several GPUs of one node: one process per GPU (torch.distributed; NCCL over NVLink/NVSwitch on the GPU box, gloo in
the CPU tests), SURVEY.md section 8(e):

  1. BGZF blocks are sharded contiguously, rank by rank (records do not straddle blocks); the BAM header lies in rank 0's range.
  2. every rank inflates + parses its shard; all-gather of {records, CB-valid reads} gives each rank the global ordinal of its
     first CB-valid read, i.e. its position in the reference's single MT19937 draw sequence (src/bam2db_ds.c:385).
  3. depth sampling on device with stream index d0 + base + local ordinal.
  4. local sort + unique of the kept keys, partitioned by an order-preserving hash of the cell (equal ranges of the cell index,
     fastf_unique_partition_device), exchanged with ONE variable-size all-to-all; every (cell, gene) group then lives on one rank.
  5. local sort + run-length dedup/count -> COO per rank on the device; the pieces are sent to rank 0 and concatenate in rank order.

The result on rank 0 is byte-identical to the single-GPU job (tests/test_sharded_gloo.py, tests/test_gpu_parity.py)."""
import ctypes as C

import numpy as np

from . import _lib
from . import bam2db_host as B


def shard_ranges(n_items, world):
    """contiguous, near-equal split of n_items over world ranks: list of (lo, hi)"""
    base, rem = divmod(n_items, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def _bgzf_block_size(buf, o):
    """size of the BGZF block whose header starts at byte o (RFC 1952 member with the BC extra subfield, SAMv1 4.1), or 0"""
    n = len(buf)
    if o + 18 > n or buf[o] != 0x1f or buf[o + 1] != 0x8b or buf[o + 2] != 8 or not (buf[o + 3] & 4):
        return 0
    xlen = int(buf[o + 10]) | (int(buf[o + 11]) << 8)
    p, end = o + 12, o + 12 + xlen
    if end > n:
        return 0
    while p + 4 <= end:
        slen = int(buf[p + 2]) | (int(buf[p + 3]) << 8)
        if buf[p] == 66 and buf[p + 1] == 67 and slen == 2 and p + 6 <= end:
            bsize = (int(buf[p + 4]) | (int(buf[p + 5]) << 8)) + 1
            return bsize if bsize >= xlen + 20 and o + bsize <= n else 0
        p += 4 + slen
    return 0


def find_block_start(buf, pos):
    """first BGZF block boundary at or behind byte pos: a header whose BSIZE leads to two more headers (or to the end of the file).
    Every rank finds its own shard this way and walks nothing but its own byte range."""
    n = len(buf)
    if pos <= 0:
        return 0
    o = pos
    while o + 18 <= n:
        # next candidate: the gzip magic with FEXTRA
        win = bytes(buf[o:min(n, o + (1 << 20))])
        k = win.find(b"\x1f\x8b\x08\x04")
        if k < 0:
            o += max(1, len(win) - 3)
            continue
        o += k
        q, good = o, 0
        while good < 3:
            bs = _bgzf_block_size(buf, q)
            if not bs:
                break
            q += bs
            good += 1
            if q == n:
                good = 3
        if good == 3:
            return o
        o += 1
    return n


def shard_bytes(buf, rank, world):
    """this rank's byte range [lo, hi) of the file: whole BGZF blocks, contiguous, rank by rank"""
    n = len(buf)
    lo = find_block_start(buf, (n * rank) // world)
    hi = find_block_start(buf, (n * (rank + 1)) // world) if rank + 1 < world else n
    return lo, max(lo, hi)


def agree(dist, torch, device, code):
    """max of an error code over all ranks (0 = everybody is fine): a rank that failed must not leave the others in a collective"""
    t = torch.tensor([int(code)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())


def index_blocks(lib, buf):
    """host BGZF index of a whole file image (numpy u8): (in_off, in_len, isize) arrays"""
    cap = max(16, buf.size // 64 + 16)
    while True:
        io, il, isz = np.zeros(cap, np.uint64), np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
        used = C.c_size_t()
        nb = lib.fastf_bgzf_index_host(C.c_void_p(buf.ctypes.data), buf.size, io.ctypes.data_as(_lib.c_u64p), il.ctypes.data_as(_lib.c_u32p), isz.ctypes.data_as(_lib.c_u32p), cap, C.byref(used))
        if nb == -2:
            cap *= 4
            continue
        if nb < 0 or used.value != buf.size:
            raise _lib.FastfError("not a whole BGZF file (index stopped at byte %d of %d)" % (used.value, buf.size))
        return io[:nb].copy(), il[:nb].copy(), isz[:nb].copy()


def sharded_tail(ctx, job, dist, torch, device, rank, world, n_cells, want_rows=False):
    """steps 2-5 above for a job that has been fed this rank's shard.  Returns on rank 0 (stats, out) like Bam2dbJob.finish; None elsewhere."""
    lib = ctx.lib
    n_rec, n_cbv = job.counts()   # the caller has agreed with the other ranks that every feed succeeded
    cnt = torch.tensor([n_rec, n_cbv], dtype=torch.int64, device=device)
    allc = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(allc, cnt)
    allc = [c.tolist() for c in allc]
    base = sum(c[1] for c in allc[:rank])
    err, rows, n, send_buf, pc, key_bits = None, None, 0, None, (C.c_uint64 * world)(), 0
    try:
        job.sample(base)
        sampled, valid = job.sample_counts()
        kp, n = job.kept_device()
        bits_cell, bits_gene, bits_umi = job.key_layout()
        key_bits = bits_cell + bits_gene + bits_umi
        if want_rows:
            rows = np.zeros(n, dtype=np.uint64)
            if n:
                ctx.check(lib.fastf_memcpy_d2h(ctx.h, C.c_void_p(rows.ctypes.data), C.c_void_p(kp), n * 8), "d2h rows")
        # 4. unique + partition by destination (order-preserving range partition of the cell index), then one all-to-all
        send_buf = torch.empty(max(n, 1), dtype=torch.int64, device=device)
        ctx.check(lib.fastf_unique_partition_device(ctx.h, C.c_void_p(kp), n, key_bits, bits_gene, bits_umi, n_cells, world, C.c_void_p(send_buf.data_ptr()), pc), "unique_partition")
    except _lib.FastfError as e:
        err = e
    if agree(dist, torch, device, 1 if err else 0):
        raise err or _lib.FastfError("another rank failed while sampling")
    send = [int(x) for x in pc]
    sc = torch.tensor(send, dtype=torch.int64, device=device)
    rc_t = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(rc_t, sc)
    recv = rc_t.tolist()
    m = sum(recv)
    recv_buf = torch.empty(max(m, 1), dtype=torch.int64, device=device)
    dist.all_to_all_single(recv_buf[:m], send_buf[:sum(send)], output_split_sizes=recv, input_split_sizes=send)
    if device != "cpu":
        torch.cuda.current_stream().synchronize()   # the library launches on its own stream
    # 5. local sort + dedup/count, results stay on the device
    nnz = C.c_uint64()
    coo_dev = torch.empty((3, max(m, 1)), dtype=torch.int32, device=device)
    try:
        if m:
            ctx.check(lib.fastf_sort_u64_device(ctx.h, C.c_void_p(recv_buf.data_ptr()), None, m, key_bits), "sort")
        ctx.check(lib.fastf_dedup_count_device_out(ctx.h, C.c_void_p(recv_buf.data_ptr()), m, bits_gene, bits_umi, C.byref(nnz), C.c_void_p(coo_dev[0].data_ptr()),
                                                   C.c_void_p(coo_dev[1].data_ptr()), C.c_void_p(coo_dev[2].data_ptr())), "dedup_count")
    except _lib.FastfError as e:
        err = e
    if agree(dist, torch, device, 1 if err else 0):
        raise err or _lib.FastfError("another rank failed while counting")
    k = nnz.value
    # 6. rank 0 collects counters and the COO pieces; pieces are ordered by rank (cells are range partitioned), so they concatenate
    meta = torch.tensor([n_rec, n_cbv, sampled, valid, k, sum(send), n], dtype=torch.int64, device=device)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    metas = [x.tolist() for x in metas]
    ks = [x[4] for x in metas]
    pieces = []
    for col in range(3):
        out_sizes = ks if rank == 0 else [0] * world
        in_sizes = [k] + [0] * (world - 1)
        dst = torch.empty(max(sum(out_sizes), 1), dtype=torch.int32, device=device)
        dist.all_to_all_single(dst[:sum(out_sizes)], coo_dev[col][:k].contiguous(), output_split_sizes=out_sizes, input_split_sizes=in_sizes)
        pieces.append(dst[:sum(out_sizes)])
    rows_all = None
    if want_rows:
        ns = [x[6] for x in metas]
        out_sizes = ns if rank == 0 else [0] * world
        dst = torch.empty(max(sum(out_sizes), 1), dtype=torch.int64, device=device)
        src = torch.empty(max(n, 1), dtype=torch.int64, device=device)
        if n:
            ctx.check(lib.fastf_memcpy_h2d(ctx.h, C.c_void_p(src.data_ptr()), C.c_void_p(rows.ctypes.data), n * 8) if device != "cpu" else 0, "rows")
            if device == "cpu":
                C.memmove(src.data_ptr(), rows.ctypes.data, n * 8)
        dist.all_to_all_single(dst[:sum(out_sizes)], src[:n], output_split_sizes=out_sizes, input_split_sizes=[n] + [0] * (world - 1))
        rows_all = dst[:sum(out_sizes)]
    if rank != 0:
        return None
    # rank 0 brings the gathered columns to the host: through the library's pinned bounce buffers (a torch .cpu() into pageable
    # memory runs at a few GB/s, and the COO of N ranks' distinct cells is N times the single-GPU one)
    def to_host(t, dtype):
        if device == "cpu":
            return t.numpy().view(dtype).copy()
        a = np.empty(t.numel(), dtype=dtype)
        if t.numel():
            ctx.check(lib.fastf_memcpy_d2h(ctx.h, C.c_void_p(a.ctypes.data), C.c_void_p(t.data_ptr()), a.nbytes), "d2h of the gathered result")
        return a
    if device != "cpu":
        torch.cuda.current_stream().synchronize()   # the collectives are ordered on torch's stream, the library copies on its own
    g, c, v = (to_host(p, np.uint32) for p in pieces)
    stats = {"total": sum(x[0] for x in metas), "cb_valid": sum(x[1] for x in metas), "sampled": sum(x[2] for x in metas), "valid": sum(x[3] for x in metas),
             "nnz": int(g.size), "bits_cell": bits_cell, "bits_gene": bits_gene, "bits_umi": bits_umi, "umi_max_bytes": (bits_umi - 4) // 8,
             "exchanged_keys": sum(x[5] for x in metas)}
    out = {"m_gene": g, "m_cell": c, "m_count": v, "row_keys": to_host(rows_all, np.uint64) if want_rows else np.zeros(0, np.uint64)}
    return stats, out


def bam2db_sharded(bam_file, db_file, path_out, barcodes_file, features_file, rate_cell, rate_depth, seed=926, dist=None, torch=None, device=None, ctx=None):
    """`fastF bam2db` over all ranks of an initialised torch.distributed group; rank 0 writes the reference's output files.
    Every rank returns 0 / 1."""
    import sys
    rank, world = dist.get_rank(), dist.get_world_size()
    own = ctx is None
    try:
        if own:
            ctx = _lib.Context(torch.cuda.current_device() if device != "cpu" else 0)
        lib = ctx.lib
        db = None
        if rank == 0:
            db = B.open_database(db_file, bam_file, barcodes_file, features_file)
            if db is None:
                raise _lib.FastfError("cannot open inputs")
            inputs = B.load_lists(db, lib, barcodes_file, features_file, rate_cell, seed)
            print("Start to convert bam file to sqlite3 database...")
            sys.stdout.flush()
        else:
            inputs = B.Bam2dbInputs(lib, barcodes_file, features_file, rate_cell, seed)
        buf = np.memmap(bam_file, dtype=np.uint8, mode="r")   # the page cache is shared: no private copy, and a rank reads nothing but its own range
        start, end = shard_bytes(buf, rank, world)
        res = None
        for umi_max_bytes in (3, 4):   # 10x UMIs are 10 or 12 bases; every rank retries together with room for 16
            code, err = 0, None
            job = None
            try:
                job = B.Bam2dbJob(ctx, inputs, rate_depth, seed, want_rows=True, umi_max_bytes=umi_max_bytes, headerless=(rank != 0))
                piece = 256 << 20
                for o in range(start, end, piece):
                    chunk = np.ascontiguousarray(buf[o:min(end, o + piece)])
                    job.feed(chunk.ctypes.data, chunk.size)
                job.counts()
            except _lib.FastfError as e:
                err, code = e, (1 if "umi-too-long" in str(e) else 2)
            worst = agree(dist, torch, device, code)   # nobody enters a collective unless everybody got this far
            if worst == 0:
                try:
                    res = sharded_tail(ctx, job, dist, torch, device, rank, world, len(inputs.cells), want_rows=True)
                finally:
                    job.close()
                break
            if job is not None:
                job.close()
            if worst == 1 and umi_max_bytes == 3:
                continue
            raise err or _lib.FastfError("another rank failed (%s)" % ("UMI longer than 16 bases" if worst == 1 else "see its message"))
        rc = 0
        if rank == 0:
            stats, out = res
            rc = B.write_outputs(db, bam_file, path_out, inputs, rate_cell, rate_depth, stats, out, ctx=ctx)
            db.close()
        return rc
    except (_lib.FastfError, ValueError, OSError) as e:
        sys.stderr.write("bam2db (rank %d): %s\n" % (rank, e))
        return 1
    finally:
        if own and ctx is not None:
            ctx.close()
