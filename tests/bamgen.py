"""TEST INFRASTRUCTURE: a tiny BAM/BGZF writer for hand-crafted edge cases (SAMv1 4.1 BGZF, 4.2 BAM records and aux fields)."""
import random
import struct
import zlib


def aux_Z(tag, s):
    return tag.encode() + b"Z" + (s if isinstance(s, bytes) else s.encode()) + b"\0"


def aux_H(tag, s):
    return tag.encode() + b"H" + s.encode() + b"\0"


def aux_int(tag, typ, v):
    fmt = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I"}[typ]
    return tag.encode() + typ.encode() + struct.pack(fmt, v)


def aux_A(tag, ch):
    return tag.encode() + b"A" + ch.encode()


def aux_f(tag, v):
    return tag.encode() + b"f" + struct.pack("<f", v)


def aux_d(tag, v):
    return tag.encode() + b"d" + struct.pack("<d", v)


def aux_B(tag, sub, vals):
    fmt = {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}[sub]
    return tag.encode() + b"B" + sub.encode() + struct.pack("<I", len(vals)) + struct.pack("<%d%s" % (len(vals), fmt), *vals)


def record(qname, aux, l_seq=20, n_cigar=1, pos=0, rng=None):
    rng = rng or random
    name = qname.encode() + b"\0"
    cigar = b"".join(struct.pack("<I", (l_seq << 4) | 0) for _ in range(n_cigar))
    seq = bytes(rng.getrandbits(8) for _ in range((l_seq + 1) // 2))
    qual = bytes(rng.randrange(2, 40) for _ in range(l_seq))
    core = struct.pack("<iiBBHHHiiii", 0, pos, len(name), 255, 4680, n_cigar, 0, l_seq, -1, -1, 0)
    body = core + name + cigar + seq + qual + b"".join(aux)
    return struct.pack("<i", len(body)) + body


def bam_header(text=b"@HD\tVN:1.6\n", refs=((b"chr1", 248956422),)):
    h = b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(refs))
    for name, ln in refs:
        h += struct.pack("<i", len(name) + 1) + name + b"\0" + struct.pack("<i", ln)
    return h


def bgzf_block(payload, level=6, strategy=zlib.Z_DEFAULT_STRATEGY):
    assert len(payload) <= 65536
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    comp = co.compress(payload) + co.flush()
    bsize = 18 + len(comp) + 8
    assert bsize <= 65536, "block does not fit: lower the payload size"
    hdr = b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", bsize - 1)
    return hdr + comp + struct.pack("<II", zlib.crc32(payload) & 0xffffffff, len(payload))


EOF_BLOCK = bgzf_block(b"")


def bgzf_file(chunks, modes=None):
    """chunks: list of byte strings, each becomes one BGZF block; modes: per chunk (level, strategy)"""
    out = b""
    for i, c in enumerate(chunks):
        level, strat = (modes[i % len(modes)] if modes else (6, zlib.Z_DEFAULT_STRATEGY))
        out += bgzf_block(c, level, strat)
    return out + EOF_BLOCK


def pack_records(header, records, max_payload=0xff00, header_own_blocks=True):
    """htslib-style blocking: the header in its own block(s), records never split across blocks"""
    chunks = []
    if header_own_blocks:
        for i in range(0, len(header), max_payload):
            chunks.append(header[i:i + max_payload])
        cur = b""
    else:
        cur = header
    for r in records:
        if len(cur) + len(r) > max_payload and cur:
            chunks.append(cur)
            cur = b""
        cur += r
    if cur:
        chunks.append(cur)
    return chunks


def edge_case_bam(barcodes, genes, seed=7, n_plain=400, umi_len=12):
    """A BAM exercising every branch the reference's loop has (and none of the inputs that crash it):
    barcodes/genes are lists of str.  Returns the BGZF bytes."""
    rng = random.Random(seed)
    B = "ACGT"

    def umi(n=umi_len):
        return "".join(rng.choice(B) for _ in range(n))

    def std(cb, gx, xf=("C", 25), ub=None, extra_front=(), extra_back=()):
        aux = list(extra_front)
        aux += [aux_int("NH", "C", 1), aux_int("HI", "C", 1), aux_int("AS", "C", 89), aux_int("nM", "C", 0)]
        if gx is not None:
            aux += [aux_Z("GX", gx), aux_Z("GN", "G" + gx[-4:])]
        aux.append(aux_int("xf", xf[0], xf[1]))
        aux += [aux_Z("CR", cb[:16]), aux_Z("CY", "F" * 16), aux_Z("CB", cb)]
        u = ub if ub is not None else umi()
        aux += [aux_Z("UR", u), aux_Z("UY", "F" * len(u)), aux_Z("UB", u)]
        aux += list(extra_back)
        return aux

    recs = []
    n = 0

    def add(aux, **kw):
        nonlocal n
        recs.append(record("r%06d" % n, aux, rng=rng, pos=n, **kw))
        n += 1

    for i in range(n_plain):                       # plain reads with duplicates so the dedup has work to do
        cb = rng.choice(barcodes)
        gx = rng.choice(genes[:20])
        add(std(cb, gx, ub=rng.choice(["ACGTACGTACGT", "TTTTGGGGCCCC", "AAAAAAAAAAAA", umi()])))
    cb0, gx0 = barcodes[0], genes[0]
    for typ, v in (("c", 25), ("C", 17), ("s", 25), ("S", 17), ("i", 25), ("I", 17), ("c", -25), ("s", 25 + 256), ("i", 25 + 65536), ("C", 0), ("C", 24), ("I", 4294967295)):
        add(std(cb0, gx0, xf=(typ, v)))            # every integer width; values that only match after truncation must NOT match
    add([aux_int("NH", "C", 1), aux_Z("GX", gx0), aux_A("xf", "x"), aux_Z("CB", cb0), aux_Z("UB", umi())])      # xf of a non-integer type -> 0
    add([aux_int("NH", "C", 1), aux_Z("GX", gx0), aux_f("xf", 25.0), aux_Z("CB", cb0), aux_Z("UB", umi())])     # float xf -> 0
    add([aux_int("NH", "C", 1), aux_int("xf", "C", 0), aux_Z("UB", umi())])                                     # no CB
    add(std("ACGTACGTACGTACGT-9", gx0))                                                                         # CB not in the list
    add(std(cb0[:-1], gx0))                                                                                     # prefix of a barcode
    add(std(cb0 + "X", gx0))                                                                                    # barcode + suffix
    add([aux_int("NH", "C", 1), aux_A("CB", "A"), aux_int("xf", "C", 25), aux_Z("GX", gx0), aux_Z("UB", umi())])  # CB of a non-string type
    add(std(cb0, gx0 + ";" + genes[1]))                                                                         # multi-gene GX: not in the table
    add(std(cb0, "ENSG_not_there"))
    add(std(cb0, gx0, ub="ACGTNACGTACG"))                                                                       # N in the UMI -> SQL NULL, still a valid row
    add(std(cb0, gx0, ub="ACGTNACGTACG"))
    add(std(cb0, genes[2], ub="NNNNNNNNNNNN"))                                                                  # group with only NULLs -> matrix row with count 0
    add(std(cb0, gx0, ub="acgtacgtacgt"))                                                                       # lower case is not ACGT
    add(std(cb0, gx0, ub="ACGTACGTAC"))                                                                         # 10-base UMI
    add(std(cb0, gx0, ub="ACGTACGTACAA"))                                                                       # same first 10 bases + AA: distinct blob
    add(std(cb0, gx0, ub="ACGTACGT"))                                                                           # 2-byte blob
    add(std(cb0, gx0, ub=""))                                                                                   # empty UMI: empty blob (non-NULL)
    add(std(cb0, gx0, ub="A"))
    a = std(cb0, gx0)
    a = [x for x in a if not x.startswith(b"UB")]
    add(a)                                                                                                      # UB missing
    add(std(cb0, gx0, extra_front=[aux_Z("CB", "ACGTACGTACGTACGT-9")]))                                         # duplicate CB: the first (not in list) wins
    add(std(barcodes[1], gx0, extra_back=[aux_Z("CB", "ACGTACGTACGTACGT-9"), aux_int("xf", "C", 0)]))           # later duplicates are ignored
    add(std(cb0, gx0, extra_front=[aux_B("ZB", "S", [1, 2, 3, 65535]), aux_B("ZC", "c", []), aux_B("ZF", "f", [1.5, 2.5]), aux_d("ZD", 3.25), aux_f("ZE", 1.0),
                                   aux_int("ZS", "s", -3), aux_int("ZI", "i", -70000), aux_H("ZH", "1AE301"), aux_A("ZA", "Q")]))   # every aux type in front of the tags
    add(std(cb0, gx0, extra_front=[aux_Z("ZL", "x" * 300)]))                                                    # a long string to step over
    add(std(cb0, gx0), l_seq=151, n_cigar=3)
    add(std(cb0, gx0), l_seq=0, n_cigar=0)
    add([])                                                                                                     # a record without any aux data
    for i in range(200):
        cb = rng.choice(barcodes + ["ACGTACGTACGTACGT-9"])
        r = rng.random()
        if r < 0.1:
            add([aux_int("NH", "C", 1), aux_int("xf", "C", 0), aux_Z("CB", cb), aux_Z("UB", umi())])           # no GX with xf 0 (fine for the reference)
        else:
            add(std(cb, rng.choice(genes), xf=("C", rng.choice([25, 25, 17, 0, 19])) if r > 0.3 else ("C", 25)))
    return recs
