// K1 -- BGZF / raw-DEFLATE inflate on device.
//
// Replaces what htslib's BGZF reader + zlib do underneath the reference's sam_open/sam_read1
// (reference src/bam2db_ds.c:141,340,360) and zlib's gzgets for FASTQ (src/filter.c:22-27).
//
// Mapping: one BGZF block (<= 64 KiB inflated, independent deflate stream) per group of G lanes;
// G = 32 gives one block per warp-sized CTA.  Every lane of a group runs the same Huffman decode in
// lock-step on a replicated 64-bit bit buffer (branch-uniform, LUT reads are shared-memory
// broadcasts); the group cooperates on (a) coalesced input prefetch: each lane holds one 32-bit word
// of the current and the next input line and refills come from a shuffle, (b) LZ77 match copies and
// (c) Huffman table construction.  Tables live in shared memory: a 10-bit primary LUT for
// literal/length codes and an 8-bit LUT for distance codes (u16 entries: symbol << 4 | code length),
// with a canonical count/sorted-symbol walk for the rare longer codes.  Output goes straight to the
// contiguous inflated buffer in HBM; match sources are read back through L1/L2.
//
// DEFLATE is RFC 1951; BGZF framing is SAMv1 section 4.1.  All three block types, multiple deflate
// blocks per BGZF block and the empty EOF block are handled.
#pragma once
#include "common.cuh"

#define FASTF_INFL_LBITS 10
#define FASTF_INFL_DBITS 8

struct FastfInflTables {
    u16 lit_lut[1 << FASTF_INFL_LBITS];
    u16 dist_lut[1 << FASTF_INFL_DBITS];
    u16 lit_sorted[288];
    u16 dist_sorted[32];
    u16 lit_cnt[16];
    u16 dist_cnt[16];
    u16 first[16];    // scratch while building: first canonical code of each length
    u16 start[16];    // scratch while building: index of the first symbol of each length in sorted[]
    u8 lens[320];     // code lengths of the block being set up (literal/length then distance)
};

struct FastfInflConst {
    u16 len_base[32];
    u16 dist_base[32];
    u8 len_extra[32];
    u8 dist_extra[32];
    u8 cl_order[20];
};

__constant__ u16 FASTF_LEN_BASE[32] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258, 0, 0, 0};
__constant__ u8 FASTF_LEN_EXTRA[32] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0, 0, 0, 0};
__constant__ u16 FASTF_DIST_BASE[32] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577, 0, 0};
__constant__ u8 FASTF_DIST_EXTRA[32] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 0, 0};
__constant__ u8 FASTF_CL_ORDER[20] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15, 0};

// ---- lane-cooperative bit reader: identical (buf, nbits, widx) in every lane of the group ----
template <int G> struct FastfBitReader {
    const u32 *wbase;
    u32 max_words;   // words that may be loaded (inside the padded compressed buffer)
    u32 widx;        // next word to move into buf
    u32 cur, nxt;    // this lane's word of the current / next G-word line
    u64 buf;
    u32 nbits;
    u32 gmask, glane;
    u32 skip_bits;   // bits of the first word that precede the stream (byte misalignment)

    __device__ __forceinline__ u32 load_word(u32 i) const { return i < max_words ? __ldg(wbase + i) : 0u; }

    __device__ __forceinline__ void init(const u8 *comp, u64 comp_total, u64 byte_off, u32 gmask_, u32 glane_)
    {
        gmask = gmask_; glane = glane_;
        u64 aligned = byte_off & ~3ull;
        wbase = (const u32 *)(comp + aligned);
        u64 mw = (comp_total - aligned) >> 2;
        max_words = mw > 0xffffffffull ? 0xffffffffu : (u32)mw;
        skip_bits = (u32)(byte_off & 3ull) * 8u;
        widx = 0; buf = 0; nbits = 0;
        cur = load_word(glane);
        nxt = load_word((u32)G + glane);
        refill();
        buf >>= skip_bits; nbits -= skip_bits;
        refill();
    }
    __device__ __forceinline__ void refill()
    {
        if (nbits <= 32) {
            u32 w = __shfl_sync(gmask, cur, (int)(widx & (G - 1)), G);
            buf |= (u64)w << nbits;
            nbits += 32;
            widx++;
            if ((widx & (G - 1)) == 0) { cur = nxt; nxt = load_word(widx + (u32)G + glane); }
        }
    }
    __device__ __forceinline__ u32 peek(u32 n) const { return (u32)buf & ((1u << n) - 1u); }
    __device__ __forceinline__ void drop(u32 n) { buf >>= n; nbits -= n; }
    __device__ __forceinline__ u32 take(u32 n) { u32 v = peek(n); drop(n); return v; }
    // bytes of the stream consumed so far (rounded up to whole bytes)
    __device__ __forceinline__ u64 bytes_consumed() const { return (((u64)widx * 32u - nbits - skip_bits) + 7u) >> 3; }
};

// Build one canonical Huffman decoding table from code lengths lens[0..n).
// Returns 0 ok, non-zero for an over-subscribed or (non-trivially) incomplete code.
template <int G>
__device__ __forceinline__ u32 fastf_build_table(FastfInflTables &T, const u8 *lens, u32 n, u16 *cnt, u16 *sorted, u16 *lut, u32 tbits, u32 gmask, u32 glane)
{
    for (u32 i = glane; i < (1u << tbits); i += G) lut[i] = 0;
    u32 bad = 0;
    if (glane == 0) {
        for (u32 l = 0; l < 16; l++) cnt[l] = 0;
        for (u32 s = 0; s < n; s++) cnt[lens[s]]++;
        i32 left = 1;
        u32 used = 0;
        for (u32 l = 1; l < 16; l++) {
            left = (left << 1) - (i32)cnt[l];
            if (left < 0) bad = 1;
            used += cnt[l];
        }
        if (left > 0 && used > 1) bad = 1;    // incomplete code: only the 0/1-symbol special case is legal
        u32 code = 0, idx = 0;
        for (u32 l = 1; l < 16; l++) {
            T.first[l] = (u16)code;
            T.start[l] = (u16)idx;
            code = (code + cnt[l]) << 1;
            idx += cnt[l];
        }
        if (!bad) {
            u16 offs[16];
            for (u32 l = 1; l < 16; l++) offs[l] = T.start[l];
            for (u32 s = 0; s < n; s++) {
                u32 l = lens[s];
                if (l) sorted[offs[l]++] = (u16)s;
            }
        }
        T.first[0] = (u16)used;
    }
    bad = __shfl_sync(gmask, bad, 0, G);
    __syncwarp(gmask);
    if (bad) return 1;
    u32 used = T.first[0];
    for (u32 i = glane; i < used; i += G) {
        u32 sym = sorted[i];
        u32 l = lens[sym];
        if (l <= tbits) {
            u32 code = (u32)T.first[l] + (i - (u32)T.start[l]);
            u32 rev = __brev(code) >> (32 - l);
            u16 e = (u16)((sym << 4) | l);
            for (u32 j = rev; j < (1u << tbits); j += (1u << l)) lut[j] = e;
        }
    }
    __syncwarp(gmask);
    return 0;
}

// Decode one symbol.  Needs >= 15 valid bits in br.buf.  Returns 0xffff when no code matches.
template <int G>
__device__ __forceinline__ u32 fastf_decode_sym(FastfBitReader<G> &br, const u16 *lut, u32 tbits, const u16 *cnt, const u16 *sorted)
{
    u32 e = lut[(u32)br.buf & ((1u << tbits) - 1u)];
    u32 l = e & 15u;
    if (l) { br.drop(l); return e >> 4; }
    // canonical walk for codes longer than the LUT (or absent codes)
    u32 code = 0, first = 0, index = 0;
    u32 bits = (u32)br.buf;
    for (u32 len = 1; len < 16; len++) {
        code |= bits & 1u;
        bits >>= 1;
        u32 c = cnt[len];
        if (code < first + c) { br.drop(len); return sorted[index + (code - first)]; }
        index += c;
        first = (first + c) << 1;
        code <<= 1;
    }
    return 0xffffu;
}

// Inflate one raw deflate stream of in_len bytes at comp+in_off into out[0..isize).  Returns status bits.
template <int G>
__device__ u32 fastf_inflate_stream(const u8 *__restrict__ comp, u64 comp_total, u64 in_off, u32 in_len, u8 *__restrict__ out, u32 isize,
                                    FastfInflTables &T, const FastfInflConst &K, u32 gmask, u32 glane)
{
    FastfBitReader<G> br;
    br.init(comp, comp_total, in_off, gmask, glane);
    u32 pos = 0, err = 0;
    bool last = false;
    while (!last && !err) {
        br.refill();
        last = br.take(1) != 0;
        u32 btype = br.take(2);
        if (btype == 0) {
            // stored: skip to the byte boundary, LEN, NLEN, raw bytes
            br.drop(br.nbits & 7u);
            br.refill();
            u32 len = br.take(16);
            br.refill();
            u32 nlen = br.take(16);
            if ((len ^ nlen) != 0xffffu) { err |= FASTF_ST_BAD_STORED; break; }
            if (pos + len > isize) { err |= FASTF_ST_OUT_OVERFLOW; break; }
            // byte position of the next unread input byte
            u64 consumed_bits = (u64)br.widx * 32u - br.nbits - br.skip_bits;   // multiple of 8 here
            u64 src_off = in_off + (consumed_bits >> 3);
            if ((consumed_bits >> 3) + len > (u64)in_len) { err |= FASTF_ST_BAD_STORED; break; }
            for (u32 i = glane; i < len; i += G) out[pos + i] = comp[src_off + i];
            pos += len;
            // restart the reader after the raw bytes, keeping in_off as the origin for the overrun check
            u64 new_off = src_off + len;
            u32 done_bytes = (u32)(new_off - in_off);
            br.init(comp, comp_total, new_off, gmask, glane);
            in_off = new_off;
            in_len -= done_bytes;
            continue;
        }
        if (btype == 3) { err |= FASTF_ST_BAD_BTYPE; break; }
        u32 hlit = 288, hdist = 32;   // fixed code: all 32 five-bit distance codes exist (30, 31 are invalid when used)
        if (btype == 1) {
            for (u32 i = glane; i < 288; i += G) T.lens[i] = (u8)(i < 144 ? 8 : (i < 256 ? 9 : (i < 280 ? 7 : 8)));
            for (u32 i = glane; i < 32; i += G) T.lens[288 + i] = 5;
            __syncwarp(gmask);
        } else {
            br.refill();
            hlit = br.take(5) + 257;
            hdist = br.take(5) + 1;
            u32 hclen = br.take(4) + 4;
            if (hlit > 286 || hdist > 30) { err |= FASTF_ST_BAD_CODELENS; break; }
            for (u32 i = glane; i < 19; i += G) T.lens[i] = 0;
            __syncwarp(gmask);
            for (u32 i = 0; i < hclen; i++) {
                br.refill();
                u32 v = br.take(3);
                if (glane == 0) T.lens[K.cl_order[i]] = (u8)v;
            }
            __syncwarp(gmask);
            // code-length code: 7-bit LUT in the distance table's storage
            if (fastf_build_table<G>(T, T.lens, 19, T.dist_cnt, T.dist_sorted, T.dist_lut, 7, gmask, glane)) { err |= FASTF_ST_BAD_CODELENS; break; }
            u32 n = hlit + hdist, i = 0, prev = 0;
            while (i < n) {
                br.refill();
                u32 sym = fastf_decode_sym<G>(br, T.dist_lut, 7, T.dist_cnt, T.dist_sorted);
                u32 rep, val;
                if (sym < 16) { rep = 1; val = sym; prev = sym; }
                else if (sym == 16) { if (i == 0) { err |= FASTF_ST_BAD_CODELENS; break; } rep = 3 + br.take(2); val = prev; }
                else if (sym == 17) { rep = 3 + br.take(3); val = 0; prev = 0; }
                else if (sym == 18) { rep = 11 + br.take(7); val = 0; prev = 0; }
                else { err |= FASTF_ST_BAD_CODELENS; break; }
                if (i + rep > n) { err |= FASTF_ST_BAD_CODELENS; break; }
                for (u32 k = glane; k < rep; k += G) T.lens[i + k] = (u8)val;
                i += rep;
            }
            if (err) break;
            __syncwarp(gmask);
            if (T.lens[256] == 0) { err |= FASTF_ST_BAD_CODELENS; break; }
        }
        if (fastf_build_table<G>(T, T.lens, hlit, T.lit_cnt, T.lit_sorted, T.lit_lut, FASTF_INFL_LBITS, gmask, glane)) { err |= FASTF_ST_BAD_CODELENS; break; }
        if (fastf_build_table<G>(T, T.lens + hlit, hdist, T.dist_cnt, T.dist_sorted, T.dist_lut, FASTF_INFL_DBITS, gmask, glane)) { err |= FASTF_ST_BAD_CODELENS; break; }

        // ---- symbol loop ----
        for (;;) {
            br.refill();
            u32 sym = fastf_decode_sym<G>(br, T.lit_lut, FASTF_INFL_LBITS, T.lit_cnt, T.lit_sorted);
            if (sym < 256) {
                if (pos >= isize) { err |= FASTF_ST_OUT_OVERFLOW; break; }
                if (glane == 0) out[pos] = (u8)sym;
                pos++;
                continue;
            }
            if (sym == 256) break;
            if (sym > 285) { err |= FASTF_ST_BAD_SYMBOL; break; }
            sym -= 257;
            u32 len = (u32)K.len_base[sym] + br.take(K.len_extra[sym]);
            br.refill();
            u32 dsym = fastf_decode_sym<G>(br, T.dist_lut, FASTF_INFL_DBITS, T.dist_cnt, T.dist_sorted);
            if (dsym >= 30) { err |= FASTF_ST_BAD_SYMBOL; break; }
            u32 dist = (u32)K.dist_base[dsym] + br.take(K.dist_extra[dsym]);
            if (dist > pos) { err |= FASTF_ST_BAD_DISTANCE; break; }
            if (pos + len > isize) { err |= FASTF_ST_OUT_OVERFLOW; break; }
            __syncwarp(gmask);   // earlier stores by other lanes of the group are ordered before these loads
            const u8 *src = out + pos - dist;
            if (dist >= len) {
                for (u32 i = glane; i < len; i += G) out[pos + i] = src[i];
            } else {
                for (u32 i = glane; i < len; i += G) out[pos + i] = src[i % dist];
            }
            pos += len;
        }
    }
    if (!err) {
        if (pos != isize) err |= FASTF_ST_SIZE_MISMATCH;
        if (br.bytes_consumed() > (u64)in_len) err |= FASTF_ST_IN_OVERRUN;
    }
    return err;
}

// One BGZF block per G lanes; CTA = one warp.  in_off/in_len describe the raw deflate payload of each
// block inside comp (header and CRC32/ISIZE trailer already stripped by the host indexer); out_off is
// the block's offset in the contiguous inflated buffer.  comp must be 4-byte aligned; comp_total is
// the number of readable bytes (>= last payload end, padded to a multiple of 4).
template <int G>
__global__ void __launch_bounds__(32) fastf_bgzf_inflate_kernel(const u8 *__restrict__ comp, u64 comp_total, const u64 *__restrict__ in_off, const u32 *__restrict__ in_len,
                                                               const u64 *__restrict__ out_off, const u32 *__restrict__ isize, u32 nblocks, u8 *__restrict__ out, u32 *__restrict__ status)
{
    __shared__ FastfInflTables tabs[32 / G];
    __shared__ FastfInflConst K;
    const u32 lane = threadIdx.x;
    K.len_base[lane] = FASTF_LEN_BASE[lane];
    K.dist_base[lane] = FASTF_DIST_BASE[lane];
    K.len_extra[lane] = FASTF_LEN_EXTRA[lane];
    K.dist_extra[lane] = FASTF_DIST_EXTRA[lane];
    if (lane < 20) K.cl_order[lane] = FASTF_CL_ORDER[lane];
    __syncwarp();
    const u32 grp = lane / G, glane = lane % G;
    const u32 gmask = (u32)(((1ull << G) - 1ull) << (grp * G));
    const u32 b = blockIdx.x * (32 / G) + grp;
    if (b >= nblocks) return;
    u32 err = fastf_inflate_stream<G>(comp, comp_total, in_off[b], in_len[b], out + out_off[b], isize[b], tabs[grp], K, gmask, glane);
    if (glane == 0) status[b] = err;
}
