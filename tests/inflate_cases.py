"""Seeded random BGZF images for the inflate kernels: payloads with skewed symbol distributions (code lengths up to 15 bits: the
canonical walk behind the 9 / 7-bit tables), runs, near and far matches, under random zlib levels / strategies / memLevels (memLevel 1
= many small deflate blocks per BGZF block).  Shared by the emulator (CPU) and the GPU tests."""
import struct
import zlib

import numpy as np

EOF_BLOCK = b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0\x1b\0\x03\0\0\0\0\0\0\0\0\0"


def _block(p, level, strat, mem):
    co = zlib.compressobj(level, zlib.DEFLATED, -15, mem, strat)
    comp = co.compress(p) + co.flush()
    if 18 + len(comp) + 8 > 65536:
        return None
    return b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", 18 + len(comp) + 8 - 1) + comp + struct.pack("<II", zlib.crc32(p) & 0xffffffff, len(p))


def images(n_images, seed=20261018, blocks_per_image=6):
    """yields (bgzf image bytes, expected inflated bytes)"""
    rng = np.random.default_rng(seed)
    for _ in range(n_images):
        payloads, blocks = [], []
        for _k in range(blocks_per_image):
            n = int(rng.integers(1, 60000))
            kind = int(rng.integers(0, 5))
            if kind == 0:      # geometric: a few frequent bytes and a long tail of rare ones -> 13-15 bit codes
                p = np.minimum(rng.geometric(0.02 + 0.3 * rng.random(), n), 255).astype(np.uint8)
            elif kind == 1:    # text-like, repeats at many distances
                words = [bytes(rng.integers(65, 91, int(rng.integers(2, 12)), dtype=np.uint8)) for _ in range(200)]
                p = np.frombuffer(b" ".join(words[int(i)] for i in rng.integers(0, 200, n // 6 + 1))[:n].ljust(n, b"."), dtype=np.uint8)
            elif kind == 2:    # runs
                p = np.repeat(rng.integers(0, 256, n // 40 + 1, dtype=np.uint8), rng.integers(1, 80, n // 40 + 1))[:n].astype(np.uint8)
                if p.size < n:
                    p = np.concatenate([p, np.zeros(n - p.size, dtype=np.uint8)])
            elif kind == 3:    # uniform random (stored or nearly so)
                p = rng.integers(0, 256, n, dtype=np.uint8)
            else:              # zipf + a copy of earlier data
                p = (rng.zipf(1.3, n) % 256).astype(np.uint8)
                if n > 3000:
                    p[n // 2:n // 2 + 1000] = p[100:1100]
            p = p.tobytes()
            level = int(rng.choice([1, 4, 6, 9]))
            strat = int(rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED]))
            mem = int(rng.choice([1, 4, 8, 9]))
            b = _block(p, level, strat, mem)
            if b is None:      # did not fit a BGZF block with those settings
                p = p[:30000]
                b = _block(p, 6, zlib.Z_DEFAULT_STRATEGY, 8)
            payloads.append(p)
            blocks.append(b)
        yield b"".join(blocks) + EOF_BLOCK, b"".join(payloads)


def max_code_length_seen(n_images=4):
    """(diagnostic) longest Huffman code zlib emitted for these payloads is not observable from Python; kept for documentation"""
    return None
