"""CPU tests of the KERNEL LOGIC: the same .cu/.cuh sources compiled with g++ under tests/emu/cuda_emu.h (a cooperative SIMT
emulator: fibers per CUDA thread, warp collectives and __syncthreads as rendezvous, deadlock detection).  This is test
infrastructure; the product never loads the emulator build.  The real parity tests are the -m gpu ones."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu_lib():
    from fastf_b200 import build
    return build.build_emu()


def _run(emu_lib, names, shuffle=None, cli=False):
    env = dict(os.environ, FASTF_GPU_LIB=emu_lib)
    if cli:
        from fastf_b200 import build
        env["FASTF_EMU_CLI"] = build.build_cli(emu=True)
    if shuffle:
        env["FASTF_EMU_SHUFFLE"] = str(shuffle)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu", "run_emu_case.py")] + names, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]


def test_emu_inflate_adversarial_streams():
    """stored / fixed / dynamic / RLE / huffman-only / empty / 64 KiB / multi-stored blocks + corrupt streams, byte-for-byte vs zlib"""
    exe = os.path.join(ROOT, "tests", "emu", "_build", "emu_inflate")
    src = os.path.join(ROOT, "tests", "emu", "emu_inflate.cpp")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-O1", "-std=c++17", "-DFASTF_EMU", "-I" + os.path.join(ROOT, "tests", "emu"), "-o", exe, src, "-lz"], check=True)
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "FAIL" not in r.stdout, r.stdout[-2000:]


def test_emu_bam2db_edge_cases(emu_lib):
    _run(emu_lib, ["edge-c0.5-r0.5-s926"])


def test_emu_bam2db_edge_cases_shuffled_schedule(emu_lib):
    _run(emu_lib, ["edge-c1.0-r1.0-s926"], shuffle=7)


def test_emu_freq_ragged(emu_lib):
    _run(emu_lib, ["freq-ragged_nonl-l16-u12", "freq-ragged_trunc-l5-u3"])


@pytest.mark.parametrize("blocks", [1, 2, 5])
def test_emu_freq_streams_in_chunks(emu_lib, blocks, tmp_path, monkeypatch):
    """freq keeps only the keys resident: the text streams through in chunks of whole BGZF blocks; lines (and keys) that cross a chunk
    boundary, chunks without a newline, a carried tail shorter than the carry window -- every golden whitelist and a synthetic FASTQ
    (its BGZF blocks end at arbitrary text positions) come out as with one chunk"""
    monkeypatch.setenv("FASTF_STREAM_CHUNK_BLOCKS", str(blocks))
    _run(emu_lib, ["freq-ragged-l16-u12", "freq-ragged_nonl-l16-u12", "freq-ragged_trunc-l5-u3", "freq-ragged-l16-u0"])
    code = (
        "import sys, os\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import fastf_b200, oracle_binding, synth_binding\n"
        "O, S = oracle_binding.load(), synth_binding.load()\n"
        "d = %r\n"
        "p = S.params(n_reads=9000, n_cells=40, seed=3, p_umi_n=0.02, zlevel=1)\n"
        "fq, st = S.fastq(p, 2)\n"
        "assert st.n_blocks > 6, st.n_blocks\n"
        "open(os.path.join(d, 'R1.fastq.gz'), 'wb').write(fq)\n"
        "import gzip\n"
        "open(os.path.join(d, 'plain.fastq'), 'wb').write(gzip.decompress(fq)[:300000])\n"
        "O.freq(os.path.join(d, 'plain.fastq'), 16, 12, os.path.join(d, 'want.txt'))\n"
        "assert fastf_b200.freq(os.path.join(d, 'plain.fastq'), d, 16, 12) == 0\n"
        "assert open(os.path.join(d, 'whitelist.txt'), 'rb').read() == open(os.path.join(d, 'want.txt'), 'rb').read(), 'plain text'\n"
        "for l, u in ((16, 12), (7, 0), (16, 15)):\n"
        "    assert fastf_b200.freq(os.path.join(d, 'R1.fastq.gz'), d, l, u) == 0\n"
        "    O.freq(os.path.join(d, 'R1.fastq.gz'), l, u, os.path.join(d, 'want.txt'))\n"
        "    assert open(os.path.join(d, 'whitelist.txt'), 'rb').read() == open(os.path.join(d, 'want.txt'), 'rb').read(), (l, u)\n"
        "print('ok')\n") % (ROOT, os.path.join(ROOT, "tests"), str(tmp_path))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, FASTF_GPU_LIB=emu_lib), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-3000:]


def test_emu_c_host_cli(emu_lib):
    """the C host (fastf_b200/host: option parsing, list readers, sqlite + gz writers, -u) linked against the emulator build"""
    _run(emu_lib, ["synth4k-c0.5-r0.5-s926", "freq-ragged-l16-u0"], cli=True)


def test_emu_mt19937_jump_ahead(emu_lib):
    """jump-ahead tables (Berlekamp-Massey characteristic polynomial, x^(2^k) mod phi) + the jump kernel against the oracle's stream"""
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from fastf_b200 import _lib\nimport oracle_binding\n"
        "O = oracle_binding.load(); ctx = _lib.Context(0)\n"
        "for seed, first, n in ((926, 1, 1500), (926, 624, 1300), (5489, 1234567, 2000)):\n"
        "    out = np.zeros(n, dtype=np.uint32)\n"
        "    ctx.check(ctx.lib.fastf_mt19937_host_from(ctx.h, seed, first, n, out.ctypes.data_as(_lib.c_u32p)), 'from')\n"
        "    assert np.array_equal(out, O.mt_stream(seed, first + n)[first:]), (seed, first)\n"
    ) % (ROOT, os.path.join(ROOT, "tests"))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, FASTF_GPU_LIB=emu_lib), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]


def test_emu_unique_counts(emu_lib):
    """-u/--umicopies (reference src/bam2db_ds.c:527-530): distinct keys and their copies from the device sort + run-length heads"""
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "from fastf_b200 import _lib, bam2db_host\n"
        "ctx = _lib.Context(0)\n"
        "for n, bits, distinct in ((0, 40, 1), (1, 64, 1), (3000, 30, 7), (20000, 54, 5000), (9000, 58, 8990)):\n"
        "    rng = np.random.default_rng(n + bits)\n"
        "    pool = rng.integers(0, 2**63, distinct, dtype=np.uint64) >> np.uint64(63 - min(bits, 63))\n"
        "    keys = pool[rng.integers(0, distinct, n)] if n else np.zeros(0, np.uint64)\n"
        "    u, c = bam2db_host.unique_counts(ctx, keys, bits)\n"
        "    wu, wc = np.unique(keys, return_counts=True)\n"
        "    assert np.array_equal(u, wu) and np.array_equal(c.astype(np.int64), wc), (n, bits)\n"
    ) % (ROOT,)
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, FASTF_GPU_LIB=emu_lib), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]


def test_emu_bgzf_crc32_check(emu_lib):
    """warp-parallel CRC-32 (32 lane segments folded by GF(2) shifts) against zlib's for ragged block sizes; a flipped payload byte in
    a stored block (still valid deflate) and a wrong CRC field are reported as bgzf-crc32-mismatch"""
    code = (
        "import sys, zlib, ctypes as C, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from fastf_b200 import _lib\nimport bamgen\n"
        "ctx = _lib.Context(0)\n"
        "def inflate(img, lanes):\n"
        "    buf = np.frombuffer(img, dtype=np.uint8); out, n, ms = C.c_void_p(), C.c_size_t(), C.c_float()\n"
        "    if ctx.lib.fastf_inflate_host(ctx.h, C.c_void_p(buf.ctypes.data), buf.size, lanes, C.byref(out), C.byref(n), C.byref(ms)): return None\n"
        "    d = C.string_at(out, n.value); ctx.lib.fastf_free(out); return d\n"
        "rng = np.random.default_rng(5)\n"
        "payloads = [bytes(rng.integers(0, 256, n, dtype=np.uint8)) for n in (1, 16, 17, 511, 513, 4099, 20001)] + [b'', b'ACGT' * 16384]\n"
        "blocks = [bamgen.bgzf_block(p, 0) for p in payloads[:-1]] + [bamgen.bgzf_block(payloads[-1])]\n"
        "img = b''.join(blocks) + bamgen.EOF_BLOCK\n"
        "assert inflate(img, 0) == b''.join(payloads)\n"
        "bad = bytearray(img); bad[len(blocks[0]) + len(blocks[1]) + len(blocks[2]) + 18 + 5 + 100] ^= 4\n"
        "assert inflate(bytes(bad), 0) is None and b'crc32' in ctx.lib.fastf_last_error(ctx.h)\n"
        "assert inflate(bytes(bad), 0x200) is not None\n"
        "bad = bytearray(img); bad[len(img) - len(bamgen.EOF_BLOCK) - 8] ^= 1\n"
        "assert inflate(bytes(bad), 32) is None and b'crc32' in ctx.lib.fastf_last_error(ctx.h)\n"
    ) % (ROOT, os.path.join(ROOT, "tests"))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, FASTF_GPU_LIB=emu_lib), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:]


def test_emu_crb_extract_golden(emu_lib):
    """`crb` and `extract` (fastf_taghist_gpu + host pre-order) against the outputs of the unmodified reference (tests/golden/tags)"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tags_cases.py")], env=dict(os.environ, FASTF_GPU_LIB=emu_lib), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.count("ok ") == 15 and r.stdout.count("ok-streamed-collided") == 9, r.stdout[-2000:]


def test_emu_c_host_cli_crb_extract(emu_lib, tmp_path):
    """the C host's crb / extract subcommands (option parsing, BST pre-order, gz / csv writers, stdout lines) on the emulator build"""
    import gzip
    from fastf_b200 import build
    cli = build.build_cli(emu=True)
    g = os.path.join(ROOT, "tests", "golden", "tags")
    r = subprocess.run([cli, "crb", "-b", os.path.join(g, "tags.bam"), "--out", "o.gz"], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.splitlines() == ["Processed all 904 reads", "Writing to file...", "Done."], r.stdout
    assert gzip.open(tmp_path / "o.gz", "rb").read() == gzip.open(os.path.join(g, "expect_tags_crb.txt.gz"), "rb").read()
    r = subprocess.run([cli, "extract", "-b", os.path.join(g, "tags.bam"), "-t", "AS", "-T", "1"], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.splitlines() == ["Processed all 1808 reads", "Valid reads: 900"], r.stdout
    assert open(tmp_path / "tag_summary.csv", "rb").read() == gzip.open(os.path.join(g, "expect_tags_extract_AS_1.csv.gz"), "rb").read()
    # where the reference dereferences NULL (string extraction of an integer tag) this build refuses
    r = subprocess.run([cli, "extract", "-b", os.path.join(g, "tags.bam"), "-t", "NH"], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 1 and "tag-not-a-string" in r.stdout


def test_emu_bam2db_streaming_is_invariant(emu_lib):
    """feeding the BAM in odd pieces and cutting chunks of different sizes (pending blocks carried across feed calls, host bytes
    staged into the ring of compressed buffers piece by piece, partial BGZF blocks carried over) never changes the result"""
    code = (
        "import sys, os, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from fastf_b200 import _lib, bam2db_host as B\n"
        "ctx = _lib.Context(0)\n"
        "for d in ('edge', 'synth4k'):\n"
        "    g = os.path.join(%r, 'tests', 'golden', d)\n"
        "    inputs = B.Bam2dbInputs(ctx.lib, os.path.join(g, 'barcodes.tsv.gz'), os.path.join(g, 'features.tsv.gz'), 0.5, 926)\n"
        "    bam = np.fromfile(os.path.join(g, 'in.bam'), dtype=np.uint8)\n"
        "    ref = None\n"
        "    for chunk, piece in ((0, 0), (1 << 20, 0), (1 << 20, 777), (3 << 20, 50001), (0, 4099), (70000, 13)):\n"
        "        if piece == 13 and d == 'synth4k': continue\n"
        "        st, out = B.run_device(ctx, bam, inputs, 0.5, 926, want_rows=True, chunk_inflated_bytes=chunk, feed_piece=piece)\n"
        "        key = (st['total'], st['cb_valid'], st['sampled'], st['valid'], st['nnz'], out['m_gene'].tobytes(), out['m_cell'].tobytes(), out['m_count'].tobytes(), out['row_keys'].tobytes())\n"
        "        if ref is None: ref = key\n"
        "        assert key == ref, (d, chunk, piece, st)\n"
        "    print('ok', d, ref[:5])\n"
    ) % (ROOT, os.path.join(ROOT, "tests"), ROOT)
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, FASTF_GPU_LIB=emu_lib), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    assert r.returncode == 0 and r.stdout.count("ok ") == 2, r.stdout[-3000:]


def test_emu_records_straddling_bgzf_blocks(emu_lib, tmp_path):
    """BAM files whose records cross BGZF block boundaries (writers other than htslib): the golden BAMs re-cut into blocks of odd
    sizes (down to 997 bytes: records spanning several blocks, blocks without any record start, a header spanning 70 blocks) must
    give the reference's recorded outputs through the operators, which retry with FASTF_BAM_STRADDLE on their own; the strict
    default still refuses such a file; the flag changes nothing for an htslib-style file"""
    code = (
        "import sys, os, gzip, json, shutil, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import fastf_b200, bamgen\n"
        "from fastf_b200 import _lib, bam2db_host as B, tags_host as T\n"
        "from dbdigest import db_digest\n"
        "ROOT, tmp = %r, %r\n"
        "ctx = _lib.Context(0)\n"
        "for d, exp, rc, rd, seed, cut in (('edge', 'expect_c0.5_r0.5_s926', 0.5, 0.5, 926, 997), ('synth4k', 'expect_c1.0_r0.3_s926', 1.0, 0.3, 926, 30011)):\n"
        "    g = os.path.join(ROOT, 'tests', 'golden', d)\n"
        "    w = os.path.join(tmp, d + str(cut)); os.makedirs(os.path.join(w, 'out'))\n"
        "    whole = gzip.decompress(open(os.path.join(g, 'in.bam'), 'rb').read())\n"
        "    open(os.path.join(w, 'in.bam'), 'wb').write(bamgen.bgzf_file([whole[i:i + cut] for i in range(0, len(whole), cut)]))\n"
        "    for f in ('barcodes.tsv.gz', 'features.tsv.gz'): shutil.copy(os.path.join(g, f), w)\n"
        "    inputs = B.Bam2dbInputs(ctx.lib, os.path.join(w, 'barcodes.tsv.gz'), os.path.join(w, 'features.tsv.gz'), rc, seed)\n"
        "    try:\n"
        "        B.run_device(ctx, np.fromfile(os.path.join(w, 'in.bam'), dtype=np.uint8), inputs, rd, seed)\n"
        "        raise SystemExit('the strict default accepted straddling records')\n"
        "    except _lib.FastfError as e:\n"
        "        assert 'straddles' in str(e), e\n"
        "    os.chdir(w)\n"
        "    B._umi_copies_flag = 1 if d == 'synth4k' else 0\n"
        "    assert fastf_b200.bam2db('in.bam', os.path.join(w, 'x.db'), os.path.join(w, 'out'), 'barcodes.tsv.gz', 'features.tsv.gz', rc, rd, seed, ctx=ctx) == 0\n"
        "    for f in ['matrix.mtx.gz', 'barcodes.tsv.gz', 'features.tsv.gz'] + (['umi.tsv.gz'] if d == 'synth4k' else []):\n"
        "        assert gzip.open(os.path.join(w, 'out', f), 'rb').read() == gzip.open(os.path.join(g, exp, f), 'rb').read(), (d, f)\n"
        "    assert db_digest(os.path.join(w, 'x.db')) == json.load(open(os.path.join(g, exp, 'db_digest.json'))), d\n"
        "    print('ok bam2db', d)\n"
        "g = os.path.join(ROOT, 'tests', 'golden', 'tags')\n"
        "whole = gzip.decompress(open(os.path.join(g, 'tags.bam'), 'rb').read())\n"
        "p = os.path.join(tmp, 'tags_cut.bam')\n"
        "open(p, 'wb').write(bamgen.bgzf_file([whole[i:i + 5003] for i in range(0, len(whole), 5003)]))\n"
        "assert T.crb(ctx, p, os.path.join(tmp, 'crb.gz')) == 904\n"
        "assert gzip.open(os.path.join(tmp, 'crb.gz'), 'rb').read() == gzip.open(os.path.join(g, 'expect_tags_crb.txt.gz'), 'rb').read()\n"
        "assert T.extract_bam(ctx, p, 'AS', 1, tmp) == (1808, 900)\n"
        "assert open(os.path.join(tmp, 'tag_summary.csv'), 'rb').read() == gzip.open(os.path.join(g, 'expect_tags_extract_AS_1.csv.gz'), 'rb').read()\n"
        "print('ok tags')\n"
    ) % (ROOT, os.path.join(ROOT, "tests"), ROOT, str(tmp_path))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, FASTF_GPU_LIB=emu_lib), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    assert r.returncode == 0 and r.stdout.count("ok ") == 3, r.stdout[-3000:]
    # the C host retries the same way
    from fastf_b200 import build
    cli = build.build_cli(emu=True)
    r = subprocess.run([cli, "crb", "-b", str(tmp_path / "tags_cut.bam"), "-o", "c.gz"], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout
    import gzip
    assert gzip.open(tmp_path / "c.gz", "rb").read() == gzip.open(os.path.join(ROOT, "tests", "golden", "tags", "expect_tags_crb.txt.gz"), "rb").read()
    w = tmp_path / "edge997"
    os.makedirs(w / "out2", exist_ok=True)
    r = subprocess.run([cli, "bam2db", "-b", "in.bam", "-f", "features.tsv.gz", "-a", "barcodes.tsv.gz", "-d", "y.db", "-c", "0.5", "-r", "0.5", "-o", "out2", "-s", "926"], cwd=w,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    g = os.path.join(ROOT, "tests", "golden", "edge", "expect_c0.5_r0.5_s926")
    assert gzip.open(w / "out2" / "matrix.mtx.gz", "rb").read() == gzip.open(os.path.join(g, "matrix.mtx.gz"), "rb").read()


def test_emu_inflate_random_streams(emu_lib):
    """seeded random BGZF images (tests/inflate_cases.py: skewed distributions with code lengths up to 15 bits, runs, near and far matches,
    random zlib levels / strategies / memLevels): byte for byte vs zlib for the thread-per-stream and the lock-step kernel"""
    code = (
        "import sys, ctypes as C, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from fastf_b200 import _lib\nimport inflate_cases\n"
        "ctx = _lib.Context(0)\n"
        "def inflate(img, lanes):\n"
        "    buf = np.frombuffer(img, dtype=np.uint8); out, n, ms = C.c_void_p(), C.c_size_t(), C.c_float()\n"
        "    assert ctx.lib.fastf_inflate_host(ctx.h, C.c_void_p(buf.ctypes.data), buf.size, lanes, C.byref(out), C.byref(n), C.byref(ms)) == 0, ctx.lib.fastf_last_error(ctx.h)\n"
        "    d = C.string_at(out, n.value); ctx.lib.fastf_free(out); return d\n"
        "k = 0\n"
        "for img, want in inflate_cases.images(24):\n"
        "    for lanes in (0, 32):\n"
        "        assert inflate(img, lanes) == want, (k, lanes)\n"
        "    k += 1\n"
        "print('ok', k)\n"
    ) % (ROOT, os.path.join(ROOT, "tests"))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, FASTF_GPU_LIB=emu_lib), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    assert r.returncode == 0 and "ok 24" in r.stdout, r.stdout[-3000:]
