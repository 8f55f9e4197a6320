"""CPU tests of the host writers behind `fastF bam2db` (fastf_b200/host/sqlite_bulk.c, fast_writers.c): the direct SQLite b-tree
loader must produce a database that sqlite itself finds intact and that holds exactly the rows a row-by-row INSERT would; the
parallel gzip writer must produce a file whose decompressed bytes are the text."""
import ctypes as C
import gzip
import os
import sqlite3
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "fastf_b200", "host")


DRIVER = r"""
#include "fastf_host.h"
#include <string.h>
/* test driver: rows given as arrays, loaded / written through the multi-threaded entry points */
typedef struct { const int64_t *cell, *gene; const int32_t *blen; const uint8_t *blob; } rows_t;
static void enc(void *c, uint64_t i, uint8_t *types, unsigned *ncol, uint8_t *body, unsigned *nb)
{
    const rows_t *R = (const rows_t *)c;
    unsigned n = fastf_sqlite_int_col(R->cell[i], &types[0], body);
    n += fastf_sqlite_int_col(R->gene[i], &types[1], body + n);
    types[2] = R->blen[i] < 0 ? 0 : (uint8_t)(12 + 2 * R->blen[i]);
    if (R->blen[i] > 0) { memcpy(body + n, R->blob + 4 * i, (size_t)R->blen[i]); n += (unsigned)R->blen[i]; }
    *ncol = 3; *nb = n;
}
int drv_load(const char *db, unsigned root, uint64_t n_seq_before, uint64_t n, const int64_t *cell, const int64_t *gene, const int32_t *blen, const uint8_t *blob)
{
    rows_t R = {cell, gene, blen, blob};
    fastf_sqlite_bulk *b = fastf_sqlite_bulk_begin(db, root);
    if (!b) return 2;
    for (uint64_t i = 0; i < n_seq_before; i++) {   /* a few rows through the streaming call first: the two paths must chain */
        int64_t v[2] = {cell[i], gene[i]};
        fastf_sqlite_bulk_row(b, v, 2, 1, blen[i] < 0 ? NULL : blob + 4 * i, blen[i] < 0 ? 0 : (unsigned)blen[i]);
    }
    rows_t R2 = {cell + n_seq_before, gene + n_seq_before, blen + n_seq_before, blob + 4 * n_seq_before};
    (void)R;
    int rc = fastf_sqlite_bulk_rows_parallel(b, n - n_seq_before, enc, &R2);
    return fastf_sqlite_bulk_end(b) | rc;
}
static unsigned fmt(void *c, uint64_t i, char *p)
{
    const rows_t *R = (const rows_t *)c;
    char *q = p;
    q += fastf_fmt_i64(q, R->gene[i]); *q++ = ' ';
    q += fastf_fmt_i64(q, R->cell[i]); *q++ = ' ';
    q += fastf_fmt_i64(q, R->blen[i]); *q++ = '\n';
    return (unsigned)(q - p);
}
int drv_lines(const char *path, const char *head, uint64_t n, const int64_t *cell, const int64_t *gene, const int32_t *blen)
{
    rows_t R = {cell, gene, blen, NULL};
    return fastf_gz_write_lines_parallel(path, head, strlen(head), n, fmt, &R);
}
"""


@pytest.fixture(scope="module")
def hostw(tmp_path_factory):
    d = tmp_path_factory.mktemp("hostw")
    so = str(d / "libhostw.so")
    (d / "driver.c").write_text(DRIVER)
    subprocess.run(["gcc", "-O2", "-std=gnu11", "-Wall", "-fPIC", "-shared", "-I" + HOST, "-o", so, os.path.join(HOST, "sqlite_bulk.c"), os.path.join(HOST, "fast_writers.c"), str(d / "driver.c"),
                    "-lz", "-lpthread"], check=True)
    lib = C.CDLL(so)
    lib.drv_load.argtypes = [C.c_char_p, C.c_uint, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.drv_lines.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.fastf_sqlite_bulk_begin.restype = C.c_void_p
    lib.fastf_sqlite_bulk_begin.argtypes = [C.c_char_p, C.c_uint]
    lib.fastf_sqlite_bulk_row.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_uint, C.c_int, C.c_void_p, C.c_uint]
    lib.fastf_sqlite_bulk_row4.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.c_uint, C.c_int64]
    lib.fastf_sqlite_bulk_end.argtypes = [C.c_void_p]
    lib.fastf_gz_write_parallel.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
    return lib


def _rows(n, seed):
    """(cell, gene, blob-or-None) rows covering every integer width sqlite distinguishes, NULLs and blob lengths 0..4"""
    r = seed
    out = []
    special = [0, 1, 2, 127, 128, 255, 32767, 32768, 8388607, 8388608, 2147483647, -1, -128, -129, -32768, -32769, 2**40, -2**40, 2**62]
    for i in range(n):
        r = (r * 6364136223846793005 + 1442695040888963407) & (2**64 - 1)
        cell = special[(r >> 8) % len(special)] if (r >> 3) % 11 == 0 else (r >> 20) % 50000 + 1
        gene = (r >> 34) % 40000 + 1
        k = (r >> 50) % 7
        blob = None if k == 5 else bytes(((r >> (8 * j)) & 0xff) for j in range(min(k, 4)))
        out.append((cell, gene, blob))
    return out


def _root(db, name):
    return db.execute("select rootpage from sqlite_master where name=?", (name,)).fetchone()[0]


@pytest.mark.parametrize("n,page_size", [(0, 4096), (1, 4096), (300, 4096), (5000, 4096), (200_000, 4096), (60_000, 512), (400_000, 512), (3000, 65536)])
def test_sqlite_bulk_loader_equals_row_inserts(hostw, tmp_path, n, page_size):
    rows = _rows(n, 7 + n)
    a, b = str(tmp_path / "bulk.db"), str(tmp_path / "plain.db")
    for path in (a, b):
        db = sqlite3.connect(path)
        db.execute("PRAGMA page_size=%d" % page_size)
        db.execute("CREATE TABLE cell (cell_barcode TEXT);")
        db.executemany("INSERT INTO cell VALUES (?)", [("AAAC%06d" % i,) for i in range(1000)])
        db.execute("CREATE TABLE umi (cell_index INTEGER, feature_index INTEGER, encoded_umi TEXT);")
        db.execute("CREATE TABLE mtx(\n  feature_index INT,\n  cell_index INT,\n  expression_level\n)")
        db.execute("CREATE TABLE numi(\n  feature_index INT,\n  cell_index INT,\n  encoded_umi TEXT,\n  n_copy\n)")
        db.commit()
        if path == b:
            db.executemany("INSERT INTO umi VALUES (?,?,?)", rows)
            db.executemany("INSERT INTO mtx VALUES (?,?,?)", [(g, c, (c * 7 + g) % 300) for c, g, _ in rows[: n // 2]])
            db.executemany("INSERT INTO numi VALUES (?,?,?,?)", [(g, c, u, (c + g) % 5 + 1) for c, g, u in rows[: n // 3]])
            db.commit()
        roots = {t: _root(db, t) for t in ("umi", "mtx", "numi")}
        db.close()
    v = (C.c_int64 * 3)()
    h = hostw.fastf_sqlite_bulk_begin(a.encode(), roots["umi"])
    assert h
    for c, g, u in rows:
        v[0], v[1] = c, g
        assert hostw.fastf_sqlite_bulk_row(h, v, 2, 1, u, len(u) if u is not None else 0) == 0
    assert hostw.fastf_sqlite_bulk_end(h) == 0
    h = hostw.fastf_sqlite_bulk_begin(a.encode(), roots["mtx"])
    assert h
    for c, g, _ in rows[: n // 2]:
        v[0], v[1], v[2] = g, c, (c * 7 + g) % 300
        assert hostw.fastf_sqlite_bulk_row(h, v, 3, 0, None, 0) == 0
    assert hostw.fastf_sqlite_bulk_end(h) == 0
    h = hostw.fastf_sqlite_bulk_begin(a.encode(), roots["numi"])
    assert h
    for c, g, u in rows[: n // 3]:
        v[0], v[1] = g, c
        assert hostw.fastf_sqlite_bulk_row4(h, v, u, len(u) if u is not None else 0, (c + g) % 5 + 1) == 0
    assert hostw.fastf_sqlite_bulk_end(h) == 0
    da, db_ = sqlite3.connect(a), sqlite3.connect(b)
    assert da.execute("PRAGMA integrity_check").fetchall() == [("ok",)]
    for t in ("cell", "umi", "mtx", "numi"):
        assert da.execute("select rowid, * from %s order by rowid" % t).fetchall() == db_.execute("select rowid, * from %s order by rowid" % t).fetchall(), t
    assert [tuple(map(type, r)) for r in da.execute("select * from umi limit 200")] == [tuple(map(type, r)) for r in db_.execute("select * from umi limit 200")]
    # the database stays an ordinary one: the reference's GROUP BY runs on it, and it can be written to again
    q = "SELECT feature_index, cell_index, COUNT(DISTINCT encoded_umi) FROM umi GROUP BY cell_index, feature_index"
    assert da.execute(q).fetchall() == db_.execute(q).fetchall()
    da.execute("INSERT INTO umi VALUES (1, 2, NULL)")
    da.execute("DELETE FROM umi WHERE rowid % 3 = 0")
    da.commit()
    assert da.execute("PRAGMA integrity_check").fetchall() == [("ok",)]
    assert da.execute("select count(*) from umi").fetchone()[0] == n + 1 - (n + 1) // 3


def test_sqlite_bulk_loader_refuses_a_table_with_rows(hostw, tmp_path):
    p = str(tmp_path / "x.db")
    db = sqlite3.connect(p)
    db.execute("CREATE TABLE umi (cell_index INTEGER, feature_index INTEGER, encoded_umi TEXT);")
    db.execute("INSERT INTO umi VALUES (1,1,NULL)")
    db.commit()
    root = _root(db, "umi")
    db.close()
    assert not hostw.fastf_sqlite_bulk_begin(p.encode(), root)
    assert not hostw.fastf_sqlite_bulk_begin(str(tmp_path / "missing.db").encode(), 2)


@pytest.mark.parametrize("n", [0, 1, 1000, (4 << 20) - 1, (4 << 20), (9 << 20) + 12345])
def test_parallel_gzip_writer_round_trips(hostw, tmp_path, n):
    text = (b"%d %d %d\n" % (123456, 7890, 42)) * (n // 15 + 1)
    text = text[:n]
    p = str(tmp_path / "t.gz")
    assert hostw.fastf_gz_write_parallel(p.encode(), text, len(text)) == 0
    assert gzip.open(p, "rb").read() == text
    assert subprocess.run(["gzip", "-t", p]).returncode == 0


@pytest.mark.parametrize("n,n_seq,page_size,threads", [(150_000, 0, 4096, 5), (150_000, 777, 4096, 3), (300_000, 10, 512, 8), (39_999, 5, 4096, 8), (1_200_000, 0, 4096, 0)])
def test_sqlite_parallel_loader_equals_row_inserts(hostw, tmp_path, monkeypatch, n, n_seq, page_size, threads):
    """leaves laid out by several threads (rows_parallel), chained behind rows from the streaming call"""
    import numpy as np
    if threads:
        monkeypatch.setenv("FASTF_HOST_THREADS", str(threads))
    rows = _rows(n, 99 + n)
    cell = np.array([r[0] for r in rows], dtype=np.int64)
    gene = np.array([r[1] for r in rows], dtype=np.int64)
    blen = np.array([-1 if r[2] is None else len(r[2]) for r in rows], dtype=np.int32)
    blob = np.zeros((n, 4), dtype=np.uint8)
    for i, r in enumerate(rows):
        if r[2]:
            blob[i, : len(r[2])] = list(r[2])
    a, b = str(tmp_path / "bulk.db"), str(tmp_path / "plain.db")
    for path in (a, b):
        db = sqlite3.connect(path)
        db.execute("PRAGMA page_size=%d" % page_size)
        db.execute("CREATE TABLE umi (cell_index INTEGER, feature_index INTEGER, encoded_umi TEXT);")
        db.commit()
        if path == b:
            db.executemany("INSERT INTO umi VALUES (?,?,?)", rows)
            db.commit()
        root = _root(db, "umi")
        db.close()
    assert hostw.drv_load(a.encode(), root, n_seq, n, cell.ctypes.data, gene.ctypes.data, blen.ctypes.data, blob.ctypes.data) == 0
    da, db_ = sqlite3.connect(a), sqlite3.connect(b)
    assert da.execute("PRAGMA integrity_check").fetchall() == [("ok",)]
    assert da.execute("select rowid, * from umi order by rowid").fetchall() == db_.execute("select rowid, * from umi order by rowid").fetchall()
    p = str(tmp_path / "lines.gz")
    head = "%%MatrixMarket test\n"
    assert hostw.drv_lines(p.encode(), head.encode(), n, cell.ctypes.data, gene.ctypes.data, blen.ctypes.data) == 0
    want = head.encode() + b"".join(b"%d %d %d\n" % (g, c, l) for c, g, l in zip(cell.tolist(), gene.tolist(), blen.tolist()))
    assert gzip.open(p, "rb").read() == want
