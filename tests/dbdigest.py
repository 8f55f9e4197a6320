"""schema text + sha256 of every table's rows (rowid order) of a bam2db sqlite file -- same function scripts/make_golden.py used on the reference's database"""
import hashlib
import sqlite3


def db_digest(path):
    c = sqlite3.connect(path)
    out = {"schema": [list(r) for r in c.execute("select name, sql from sqlite_master order by name")]}
    for (name,) in c.execute("select name from sqlite_master where type='table' order by name").fetchall():
        h = hashlib.sha256()
        n = 0
        for row in c.execute("select * from %s order by rowid" % name):
            h.update(repr(row).encode())
            n += 1
        out[name] = [n, h.hexdigest()]
    c.close()
    return out
