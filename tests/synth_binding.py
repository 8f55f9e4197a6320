"""ctypes binding of the synthetic 10x-v3 BAM / FASTQ generator (fastf_b200/synth)."""
import ctypes as C
import gzip
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class Params(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("n_molecules", C.c_uint64), ("seed", C.c_uint64), ("n_cells", C.c_uint32), ("n_genes", C.c_uint32),
                ("umi_len", C.c_uint32), ("zlevel", C.c_int32), ("p_cb_in_list", C.c_double), ("p_cb_not_in_list", C.c_double),
                ("p_gx25", C.c_double), ("p_gx17", C.c_double), ("p_umi_n", C.c_double), ("p_bc_error", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("n_blocks", C.c_uint64), ("inflated_bytes", C.c_uint64), ("compressed_bytes", C.c_uint64)]


class Synth:
    def __init__(self, lib):
        self.lib = lib
        u8pp = C.POINTER(C.POINTER(C.c_uint8))
        lib.fastf_synth_defaults.argtypes = [C.POINTER(Params)]
        lib.fastf_synth_bam.argtypes = [C.POINTER(Params), C.c_int, u8pp, C.POINTER(C.c_size_t), C.POINTER(Stats)]
        lib.fastf_synth_fastq.argtypes = [C.POINTER(Params), C.c_int, u8pp, C.POINTER(C.c_size_t), C.POINTER(Stats)]
        lib.fastf_synth_barcodes.argtypes = [C.POINTER(Params), C.POINTER(C.c_char_p), C.POINTER(C.c_size_t)]
        lib.fastf_synth_features.argtypes = [C.POINTER(Params), C.POINTER(C.c_char_p), C.POINTER(C.c_size_t)]
        lib.fastf_synth_free.argtypes = [C.c_void_p]

    def params(self, **kw):
        p = Params()
        self.lib.fastf_synth_defaults(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def _gen(self, fn, p, threads):
        out, n, st = C.POINTER(C.c_uint8)(), C.c_size_t(), Stats()
        if fn(C.byref(p), threads, C.byref(out), C.byref(n), C.byref(st)):
            raise RuntimeError("synth failed")
        # images beyond 2 GiB: ctypes.string_at takes a C int
        data = memoryview((C.c_char * n.value).from_address(C.addressof(out.contents))).tobytes() if n.value else b""
        self.lib.fastf_synth_free(out)
        return data, st

    def bam(self, p, threads=0):
        return self._gen(self.lib.fastf_synth_bam, p, threads or (os.cpu_count() or 1))

    def fastq(self, p, threads=0):
        return self._gen(self.lib.fastf_synth_fastq, p, threads or (os.cpu_count() or 1))

    def _text(self, fn, p):
        out, n = C.c_char_p(), C.c_size_t()
        fn(C.byref(p), C.byref(out), C.byref(n))
        data = C.string_at(out, n.value)
        return data

    def barcodes(self, p):
        return self._text(self.lib.fastf_synth_barcodes, p)

    def features(self, p):
        return self._text(self.lib.fastf_synth_features, p)

    def write_bam_set(self, d, **kw):
        """writes synth.bam, barcodes.tsv.gz, features.tsv.gz into directory d; returns (paths, stats)"""
        p = self.params(**kw)
        bam, st = self.bam(p)
        paths = {"bam": os.path.join(d, "synth.bam"), "barcodes": os.path.join(d, "barcodes.tsv.gz"), "features": os.path.join(d, "features.tsv.gz")}
        open(paths["bam"], "wb").write(bam)
        with gzip.open(paths["barcodes"], "wb") as f:
            f.write(self.barcodes(p))
        with gzip.open(paths["features"], "wb") as f:
            f.write(self.features(p))
        return paths, st

    def write_fastq(self, d, **kw):
        p = self.params(**kw)
        fq, st = self.fastq(p)
        path = os.path.join(d, "R1.fastq.gz")
        open(path, "wb").write(fq)
        return path, st


_s = None


def load():
    global _s
    if _s is None:
        from fastf_b200 import build
        _s = Synth(C.CDLL(build.build_synth()))
    return _s
