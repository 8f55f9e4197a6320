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
    n_rec, n_cbv = job.counts()
    cnt = torch.tensor([n_rec, n_cbv], dtype=torch.int64, device=device)
    allc = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(allc, cnt)
    allc = [c.tolist() for c in allc]
    base = sum(c[1] for c in allc[:rank])
    job.sample(base)
    sampled, valid = job.sample_counts()
    kp, n = job.kept_device()
    bits_cell, bits_gene, bits_umi = job.key_layout()
    key_bits = bits_cell + bits_gene + bits_umi
    rows = None
    if want_rows:
        rows = np.zeros(n, dtype=np.uint64)
        if n:
            ctx.check(lib.fastf_memcpy_d2h(ctx.h, C.c_void_p(rows.ctypes.data), C.c_void_p(kp), n * 8), "d2h rows")
    # 4. unique + partition by destination (order-preserving range partition of the cell index), then one all-to-all
    send_buf = torch.empty(max(n, 1), dtype=torch.int64, device=device)
    pc = (C.c_uint64 * world)()
    ctx.check(lib.fastf_unique_partition_device(ctx.h, C.c_void_p(kp), n, key_bits, bits_gene, bits_umi, n_cells, world, C.c_void_p(send_buf.data_ptr()), pc), "unique_partition")
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
    if m:
        ctx.check(lib.fastf_sort_u64_device(ctx.h, C.c_void_p(recv_buf.data_ptr()), None, m, key_bits), "sort")
    ctx.check(lib.fastf_dedup_count_device_out(ctx.h, C.c_void_p(recv_buf.data_ptr()), m, bits_gene, bits_umi, C.byref(nnz), C.c_void_p(coo_dev[0].data_ptr()),
                                               C.c_void_p(coo_dev[1].data_ptr()), C.c_void_p(coo_dev[2].data_ptr())), "dedup_count")
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
    g, c, v = (p.cpu().numpy().view(np.uint32) for p in pieces)
    stats = {"total": sum(x[0] for x in metas), "cb_valid": sum(x[1] for x in metas), "sampled": sum(x[2] for x in metas), "valid": sum(x[3] for x in metas),
             "nnz": int(g.size), "bits_cell": bits_cell, "bits_gene": bits_gene, "bits_umi": bits_umi, "umi_max_bytes": (bits_umi - 4) // 8,
             "exchanged_keys": sum(x[5] for x in metas)}
    out = {"m_gene": g, "m_cell": c, "m_count": v, "row_keys": rows_all.cpu().numpy().view(np.uint64) if want_rows else np.zeros(0, np.uint64)}
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
        buf = np.fromfile(bam_file, dtype=np.uint8)   # page cache; each rank only touches its own byte range below
        io, il, isz = index_blocks(lib, buf)
        # contiguous block ranges; the BAM header must lie inside rank 0's range (the device header walk fails loudly otherwise)
        lo, hi = shard_ranges(len(io), world)[rank]
        start = int(io[lo]) - 18 if lo < len(io) else buf.size     # 18 = BGZF header with the 6-byte BC extra field
        end = int(io[hi]) - 18 if hi < len(io) else buf.size
        with B.Bam2dbJob(ctx, inputs, rate_depth, seed, want_rows=True, umi_max_bytes=3, headerless=(rank != 0)) as job:
            if end > start:
                job.feed(buf.ctypes.data + start, end - start)
            res = sharded_tail(ctx, job, dist, torch, device, rank, world, len(inputs.cells), want_rows=True)
        rc = 0
        if rank == 0:
            stats, out = res
            rc = B.write_outputs(db, bam_file, path_out, inputs, rate_cell, rate_depth, stats, out)
            db.close()
        return rc
    except (_lib.FastfError, ValueError, OSError) as e:
        sys.stderr.write("bam2db (rank %d): %s\n" % (rank, e))
        return 1
    finally:
        if own and ctx is not None:
            ctx.close()
