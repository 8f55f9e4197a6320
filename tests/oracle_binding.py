"""TEST INFRASTRUCTURE: ctypes binding of oracle/_build/liboracle.so (the CPU restatement of the reference)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
u8p, u32p, u64p, i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_int32)


class Bam2dbResult(C.Structure):
    _fields_ = [("n_cells", C.c_uint64), ("n_cells_sampled", C.c_uint64), ("n_cell_rows", C.c_uint64), ("n_features", C.c_uint64),
                ("d0", C.c_uint64), ("total", C.c_uint64), ("cb_valid", C.c_uint64), ("sampled", C.c_uint64), ("valid", C.c_uint64),
                ("n_rows", C.c_uint64), ("row_cell", u32p), ("row_gene", u32p), ("row_umi_nbytes", i32p), ("row_umi", u64p),
                ("nnz", C.c_uint64), ("m_gene", u32p), ("m_cell", u32p), ("m_count", u32p)]


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        lib.oracle_mt_stream.argtypes = [C.c_uint32, C.c_uint64, u32p]
        lib.oracle_depth_keep.argtypes = [C.c_uint32, C.c_float]
        lib.oracle_depth_keep.restype = C.c_int
        lib.oracle_sample_cells.argtypes = [C.c_uint64, C.c_float, C.c_uint32, u64p, u64p]
        lib.oracle_sample_cells.restype = C.c_uint64
        lib.oracle_encode_dna.argtypes = [C.c_char_p, u8p, C.c_size_t]
        lib.oracle_bgzf_inflate.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_bam2db.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_float, C.c_float, C.c_uint, C.c_char_p, C.POINTER(Bam2dbResult)]
        lib.oracle_bam2db_free.argtypes = [C.POINTER(Bam2dbResult)]
        lib.oracle_freq.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, C.c_char_p, u64p, u64p]
        lib.oracle_crb.argtypes = [C.c_char_p, C.c_char_p, u64p]
        lib.oracle_extract.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, u64p, u64p]

    def mt_stream(self, seed, n):
        out = np.zeros(n, dtype=np.uint32)
        self.lib.oracle_mt_stream(seed, n, out.ctypes.data_as(u32p))
        return out

    def depth_keep(self, u, rate):
        return bool(self.lib.oracle_depth_keep(int(u), C.c_float(rate)))

    def sample_cells(self, n, rate, seed):
        out = np.zeros(max(n, 1), dtype=np.uint64)
        d0 = C.c_uint64()
        ns = self.lib.oracle_sample_cells(n, C.c_float(rate), seed, out.ctypes.data_as(u64p), C.byref(d0))
        return ns, out[:ns] if ns != 2**64 - 1 else None, d0.value

    def inflate(self, data):
        buf = np.frombuffer(data, dtype=np.uint8)
        out, n = C.c_void_p(), C.c_size_t()
        rc = self.lib.oracle_bgzf_inflate(C.c_void_p(buf.ctypes.data), buf.size, C.byref(out), C.byref(n), None, None, None, None)
        if rc:
            raise RuntimeError("oracle inflate rc=%d" % rc)
        res = C.string_at(out, n.value)
        libc = C.CDLL(None)
        libc.free.argtypes = [C.c_void_p]
        libc.free(out)
        return res

    def bam2db(self, bam, barcodes, features, rate_cell, rate_depth, seed, out_dir=None):
        r = Bam2dbResult()
        rc = self.lib.oracle_bam2db(bam.encode(), barcodes.encode(), features.encode(), C.c_float(rate_cell), C.c_float(rate_depth), seed,
                                    out_dir.encode() if out_dir else None, C.byref(r))
        if rc:
            raise RuntimeError("oracle bam2db rc=%d" % rc)
        d = {k: getattr(r, k) for k in ("n_cells", "n_cells_sampled", "n_cell_rows", "n_features", "d0", "total", "cb_valid", "sampled", "valid", "n_rows", "nnz")}
        for k, n in (("row_cell", r.n_rows), ("row_gene", r.n_rows), ("row_umi_nbytes", r.n_rows), ("row_umi", r.n_rows), ("m_gene", r.nnz), ("m_cell", r.nnz), ("m_count", r.nnz)):
            d[k] = np.ctypeslib.as_array(getattr(r, k), (n,)).copy() if n else np.zeros(0)
        self.lib.oracle_bam2db_free(C.byref(r))
        return d

    def freq(self, r1, l, u, out_path):
        nr, nk = C.c_uint64(), C.c_uint64()
        rc = self.lib.oracle_freq(r1.encode(), l, u, out_path.encode(), C.byref(nr), C.byref(nk))
        if rc:
            raise RuntimeError("oracle freq rc=%d" % rc)
        return nr.value, nk.value


    def crb(self, bam, out_path):
        """-> read_count; out_path gets the decompressed text of the reference's gz output; rc 3 = the reference would dereference NULL"""
        nr = C.c_uint64()
        rc = self.lib.oracle_crb(bam.encode(), out_path.encode(), C.byref(nr))
        if rc:
            raise RuntimeError("oracle crb rc=%d" % rc)
        return nr.value

    def extract(self, bam, tag, type_, out_path):
        """-> (total_count as the reference prints it, valid_count)"""
        t, v = C.c_uint64(), C.c_uint64()
        rc = self.lib.oracle_extract(bam.encode(), tag.encode(), type_, out_path.encode(), C.byref(t), C.byref(v))
        if rc:
            raise RuntimeError("oracle extract rc=%d" % rc)
        return t.value, v.value


_o = None


def load():
    global _o
    if _o is None:
        src = os.path.join(ROOT, "oracle", "fastf_oracle.c")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"], check=True)
        _o = Oracle(C.CDLL(LIB))
    return _o
