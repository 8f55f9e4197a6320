// Host-side BGZF block indexer (SAMv1 section 4.1): walks the BSIZE chain of a byte range and
// returns, per block, where its raw-deflate payload sits and how large it inflates.  This is the
// only serial step in front of the device inflate; it touches ~20 bytes per <=64 KiB block.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <vector>

struct FastfBgzfBlock {
    uint64_t in_off;    // offset of the deflate payload relative to the indexed buffer's origin
    uint32_t in_len;    // payload bytes
    uint32_t isize;     // inflated bytes (trailer ISIZE)
    uint32_t crc32;     // trailer CRC32
};

enum { FASTF_BGZF_OK = 0, FASTF_BGZF_NEED_MORE = 1, FASTF_BGZF_BAD_MAGIC = 2, FASTF_BGZF_NO_BSIZE = 3, FASTF_BGZF_BAD_ISIZE = 4, FASTF_BGZF_LIMIT = 5 };

// Index whole blocks found in buf[0..n).  origin is added to every in_off.  *consumed = bytes covered by
// complete blocks.  Returns FASTF_BGZF_OK when the range ends on a block boundary, FASTF_BGZF_NEED_MORE when
// a trailing partial block remains, or an error code (blocks before the error are still appended).
// With max_bytes != 0 the walk stops (FASTF_BGZF_LIMIT) in front of the block that would take the run past max_bytes of
// inflated or of compressed span, or past max_blocks blocks -- one streaming chunk's worth -- so that the caller can hand that
// chunk to the device before touching the next block headers.
static inline int fastf_bgzf_index(const uint8_t *buf, size_t n, uint64_t origin, std::vector<FastfBgzfBlock> &out, size_t *consumed, uint64_t max_bytes = 0, size_t max_blocks = 0)
{
    size_t pos = 0;
    int rc = FASTF_BGZF_OK;
    uint64_t infl = 0;
    const size_t n0 = out.size();
    while (pos < n) {
        if (n - pos < 12) { rc = FASTF_BGZF_NEED_MORE; break; }
        const uint8_t *h = buf + pos;
        if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) { rc = FASTF_BGZF_BAD_MAGIC; break; }
        uint32_t xlen = (uint32_t)h[10] | ((uint32_t)h[11] << 8);
        if (n - pos < 12 + (size_t)xlen) { rc = FASTF_BGZF_NEED_MORE; break; }
        uint32_t bsize = 0;
        for (uint32_t x = 0; x + 4 <= xlen;) {
            const uint8_t *sf = h + 12 + x;
            uint32_t slen = (uint32_t)sf[2] | ((uint32_t)sf[3] << 8);
            if (sf[0] == 'B' && sf[1] == 'C' && slen == 2 && x + 6 <= xlen) bsize = ((uint32_t)sf[4] | ((uint32_t)sf[5] << 8)) + 1;
            x += 4 + slen;
        }
        if (bsize == 0 || bsize < 12 + xlen + 8) { rc = FASTF_BGZF_NO_BSIZE; break; }
        if (n - pos < bsize) { rc = FASTF_BGZF_NEED_MORE; break; }
        const uint8_t *t = h + bsize - 8;
        FastfBgzfBlock b;
        b.in_off = origin + pos + 12 + xlen;
        b.in_len = bsize - 12 - xlen - 8;
        b.crc32 = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
        b.isize = (uint32_t)t[4] | ((uint32_t)t[5] << 8) | ((uint32_t)t[6] << 16) | ((uint32_t)t[7] << 24);
        if (b.isize > 65536) { rc = FASTF_BGZF_BAD_ISIZE; break; }
        if (max_bytes && out.size() > n0 && (infl + b.isize > max_bytes || b.in_off + b.in_len - out[n0].in_off > max_bytes || out.size() - n0 >= max_blocks)) { rc = FASTF_BGZF_LIMIT; break; }
        infl += b.isize;
        out.push_back(b);
        pos += bsize;
    }
    *consumed = pos;
    return rc;
}
