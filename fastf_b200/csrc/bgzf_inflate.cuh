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

// Decode-table entry (u32), everything a symbol needs folded into one shared-memory read:
//   bits 0-3   code length in bits (0 = not in the primary table: canonical walk)
//   bits 4-7   number of extra bits that follow the code
//   bits 8-9   kind: 0 literal / plain symbol, 1 length or distance base, 2 end of block, 3 invalid symbol
//   bits 16-31 literal byte, code-length symbol, or the length / distance base value
#define FASTF_E_LIT 0u
#define FASTF_E_BASE (1u << 8)
#define FASTF_E_EOB (2u << 8)
#define FASTF_E_BAD (3u << 8)

struct FastfInflTables {
    u32 lit_lut[1 << FASTF_INFL_LBITS];
    u32 dist_lut[1 << FASTF_INFL_DBITS];
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

enum { FASTF_ALPHA_PLAIN = 0, FASTF_ALPHA_LITLEN = 1, FASTF_ALPHA_DIST = 2 };

// table entry of symbol `sym` (without the code length) for one of the three alphabets
__device__ __forceinline__ u32 fastf_make_entry(const FastfInflConst &K, u32 alpha, u32 sym)
{
    if (alpha == FASTF_ALPHA_PLAIN) return FASTF_E_LIT | (sym << 16);
    if (alpha == FASTF_ALPHA_LITLEN) {
        if (sym < 256) return FASTF_E_LIT | (sym << 16);
        if (sym == 256) return FASTF_E_EOB;
        if (sym > 285) return FASTF_E_BAD;
        return FASTF_E_BASE | ((u32)K.len_extra[sym - 257] << 4) | ((u32)K.len_base[sym - 257] << 16);
    }
    if (sym >= 30) return FASTF_E_BAD;
    return FASTF_E_BASE | ((u32)K.dist_extra[sym] << 4) | ((u32)K.dist_base[sym] << 16);
}

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
__device__ __forceinline__ u32 fastf_build_table(FastfInflTables &T, const FastfInflConst &K, u32 alpha, const u8 *lens, u32 n, u16 *cnt, u16 *sorted, u32 *lut, u32 tbits, u32 gmask, u32 glane)
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
            u32 e = fastf_make_entry(K, alpha, sym) | l;
            for (u32 j = rev; j < (1u << tbits); j += (1u << l)) lut[j] = e;
        }
    }
    __syncwarp(gmask);
    return 0;
}

// Decode one symbol without consuming it: returns its table entry (code length in bits 0-3).  Needs >= 15 valid
// bits in br.buf.  An undecodable code returns FASTF_E_BAD with length 0.
template <int G>
__device__ __forceinline__ u32 fastf_decode_entry(const FastfBitReader<G> &br, const FastfInflConst &K, u32 alpha, const u32 *lut, u32 tbits, const u16 *cnt, const u16 *sorted)
{
    u32 e = lut[(u32)br.buf & ((1u << tbits) - 1u)];
    if (e & 15u) return e;
    // canonical walk for codes longer than the primary table (or absent codes)
    u32 code = 0, first = 0, index = 0;
    u32 bits = (u32)br.buf;
    for (u32 len = 1; len < 16; len++) {
        code |= bits & 1u;
        bits >>= 1;
        u32 c = cnt[len];
        if (code < first + c) return fastf_make_entry(K, alpha, sorted[index + (code - first)]) | len;
        index += c;
        first = (first + c) << 1;
        code <<= 1;
    }
    return FASTF_E_BAD;
}

// Output side of one group: literal runs and match pieces are written as coalesced <= G-byte stores; match pieces
// are kept in registers for FASTF_PENDING matches so that the L2 round trip of their source loads overlaps the
// decoding of the following symbols.
#define FASTF_PEND_SLOTS 3
template <int G> struct FastfOutQueue {
    u8 *out;
    u32 glane;
    u32 run_start, run_n, run_byte;          // literal run: lane i holds literal i
    u32 ppos[FASTF_PEND_SLOTS], pn[FASTF_PEND_SLOTS], pbyte[FASTF_PEND_SLOTS];   // [0] = oldest

    __device__ __forceinline__ void init(u8 *o, u32 gl)
    {
        out = o; glane = gl; run_start = 0; run_n = 0; run_byte = 0;
#pragma unroll
        for (int k = 0; k < FASTF_PEND_SLOTS; k++) { ppos[k] = 0; pn[k] = 0; pbyte[k] = 0; }
    }
    __device__ __forceinline__ void literal(u32 pos, u32 byte)
    {
        if (run_n == 0) run_start = pos;
        if (glane == run_n) run_byte = byte;
        run_n++;
        if (run_n == (u32)G) flush_run();
    }
    __device__ __forceinline__ void flush_run()
    {
        if (glane < run_n) out[run_start + glane] = (u8)run_byte;
        run_n = 0;
    }
    __device__ __forceinline__ void store_slot(int k)
    {
        if (glane < pn[k]) out[ppos[k] + glane] = (u8)pbyte[k];
        pn[k] = 0;
    }
    __device__ __forceinline__ void flush_pending()
    {
#pragma unroll
        for (int k = 0; k < FASTF_PEND_SLOTS; k++) store_slot(k);
    }
    // lowest output position that is still only in registers (or ~0u)
    __device__ __forceinline__ u32 pending_lo() const
    {
        u32 lo = 0xffffffffu;
#pragma unroll
        for (int k = FASTF_PEND_SLOTS - 1; k >= 0; k--)
            if (pn[k]) lo = ppos[k];
        return lo;
    }
    __device__ __forceinline__ void push(u32 pos, u32 n, u32 byte)
    {
        store_slot(0);
#pragma unroll
        for (int k = 0; k + 1 < FASTF_PEND_SLOTS; k++) { ppos[k] = ppos[k + 1]; pn[k] = pn[k + 1]; pbyte[k] = pbyte[k + 1]; }
        ppos[FASTF_PEND_SLOTS - 1] = pos; pn[FASTF_PEND_SLOTS - 1] = n; pbyte[FASTF_PEND_SLOTS - 1] = byte;
    }
};

// Inflate one raw deflate stream of in_len bytes at comp+in_off into out[0..isize).  Returns status bits.
template <int G>
__device__ u32 fastf_inflate_stream(const u8 *__restrict__ comp, u64 comp_total, u64 in_off, u32 in_len, u8 *out, u32 isize,
                                    FastfInflTables &T, const FastfInflConst &K, u32 gmask, u32 glane)
{
    FastfBitReader<G> br;
    br.init(comp, comp_total, in_off, gmask, glane);
    FastfOutQueue<G> Q;
    Q.init(out, glane);
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
            Q.flush_run();
            Q.flush_pending();
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
            // code-length code: 7-bit table in the distance table's storage
            if (fastf_build_table<G>(T, K, FASTF_ALPHA_PLAIN, T.lens, 19, T.dist_cnt, T.dist_sorted, T.dist_lut, 7, gmask, glane)) { err |= FASTF_ST_BAD_CODELENS; break; }
            u32 n = hlit + hdist, i = 0, prev = 0;
            while (i < n) {
                br.refill();
                u32 e = fastf_decode_entry<G>(br, K, FASTF_ALPHA_PLAIN, T.dist_lut, 7, T.dist_cnt, T.dist_sorted);
                if ((e & 15u) == 0) { err |= FASTF_ST_BAD_CODELENS; break; }
                br.drop(e & 15u);
                u32 sym = e >> 16;
                u32 rep, val;
                if (sym < 16) { rep = 1; val = sym; prev = sym; }
                else if (sym == 16) { if (i == 0) { err |= FASTF_ST_BAD_CODELENS; break; } rep = 3 + br.take(2); val = prev; }
                else if (sym == 17) { rep = 3 + br.take(3); val = 0; prev = 0; }
                else { rep = 11 + br.take(7); val = 0; prev = 0; }
                if (i + rep > n) { err |= FASTF_ST_BAD_CODELENS; break; }
                for (u32 k = glane; k < rep; k += G) T.lens[i + k] = (u8)val;
                i += rep;
            }
            if (err) break;
            __syncwarp(gmask);
            if (T.lens[256] == 0) { err |= FASTF_ST_BAD_CODELENS; break; }
        }
        if (fastf_build_table<G>(T, K, FASTF_ALPHA_LITLEN, T.lens, hlit, T.lit_cnt, T.lit_sorted, T.lit_lut, FASTF_INFL_LBITS, gmask, glane)) { err |= FASTF_ST_BAD_CODELENS; break; }
        if (fastf_build_table<G>(T, K, FASTF_ALPHA_DIST, T.lens + hlit, hdist, T.dist_cnt, T.dist_sorted, T.dist_lut, FASTF_INFL_DBITS, gmask, glane)) { err |= FASTF_ST_BAD_CODELENS; break; }

        // ---- symbol loop ----
        for (;;) {
            br.refill();
            u32 e = fastf_decode_entry<G>(br, K, FASTF_ALPHA_LITLEN, T.lit_lut, FASTF_INFL_LBITS, T.lit_cnt, T.lit_sorted);
            const u32 kind = e & (3u << 8);
            if (kind == FASTF_E_LIT) {
                if (pos >= isize) { err |= FASTF_ST_OUT_OVERFLOW; break; }
                br.drop(e & 15u);
                Q.literal(pos, e >> 16);
                pos++;
                continue;
            }
            if (kind == FASTF_E_EOB) { br.drop(e & 15u); break; }
            if (kind == FASTF_E_BAD) { err |= FASTF_ST_BAD_SYMBOL; break; }
            br.drop(e & 15u);
            const u32 len = (e >> 16) + br.take((e >> 4) & 15u);
            br.refill();
            const u32 d = fastf_decode_entry<G>(br, K, FASTF_ALPHA_DIST, T.dist_lut, FASTF_INFL_DBITS, T.dist_cnt, T.dist_sorted);
            if ((d & (3u << 8)) != FASTF_E_BASE) { err |= FASTF_ST_BAD_SYMBOL; break; }
            br.drop(d & 15u);
            const u32 dist = (d >> 16) + br.take((d >> 4) & 15u);
            if (dist > pos) { err |= FASTF_ST_BAD_DISTANCE; break; }
            if (pos + len > isize) { err |= FASTF_ST_OUT_OVERFLOW; break; }
            // the copy: sources are [pos-dist, pos-dist+min(len,dist)).  Literals go out now; match pieces still held in
            // registers must be written first only when the source reaches into them.
            Q.flush_run();
            const u32 s0 = pos - dist;
            const u32 s1 = s0 + (len < dist ? len : dist);
            if (s1 > Q.pending_lo()) Q.flush_pending();
            __syncwarp(gmask);   // earlier stores by other lanes of the group are ordered before these loads
            const u8 *src = out + s0;
            for (u32 k = 0; k < len; k += G) {
                const u32 j = k + glane;
                u32 b = 0;
                if (j < len) b = src[dist >= len ? j : j % dist];
                Q.push(pos + k, (len - k) < (u32)G ? (len - k) : (u32)G, b);
            }
            pos += len;
        }
    }
    Q.flush_run();
    Q.flush_pending();
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
__global__ void __launch_bounds__(32, G == 32 ? 32 : 16) fastf_bgzf_inflate_kernel(const u8 *__restrict__ comp, u64 comp_total, const u64 *__restrict__ in_off, const u32 *__restrict__ in_len,
                                                               const u64 *__restrict__ out_off, const u32 *__restrict__ isize, u32 nblocks, u8 *out, u32 *__restrict__ status)
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
