// K6 -- per-record aux-tag extraction for the reference's other two BAM histograms:
//   `fastF crb`      (src/extract.c:64-133):  for every record with a CB tag, count the pair (CB string, CR string)
//   `fastF extract`  (src/extract.c:135-216): for every record with tag T, count its value (string, or integer printed "%d")
// Both are "group equal values, remember the first occurrence" over the records in file order -- the same shape as freq.
// A record contributes one 64-bit KEY plus the location of its value bytes in the inflated stream:
//   string mode : key = 64-bit hash of the value bytes (pair: hash of A, then B folded in); equal keys are re-checked byte for
//                 byte afterwards (fastf_taghist_verify_kernel), a collision re-runs the job with another seed -- never a
//                 silently merged group;
//   integer mode: key = the 32-bit value itself (bam_aux2i semantics: c C s S i I, every other type 0), no hashing.
// Kernel shape = fastf_bam_parse_kernel (one warp per BGZF block, 12 KB shared-memory window, lane k parses record k,
// ballot compaction in record order); records larger than the window are walked from global memory by lane 0.
#pragma once
#include "bam_parse.cuh"

#define FASTF_TAG_MODE_STRING 0u
#define FASTF_TAG_MODE_INT 1u

struct FastfTagQuery {
    u32 a0, a1;        // tag A (required for a record to count)
    u32 b0, b1;        // tag B (pair mode), b0 = 0: none
    u32 mode;          // FASTF_TAG_MODE_*
    u64 seed;          // hash seed
    u64 key_mask;      // ~0; tests AND the keys with less to force collisions
};

__device__ __forceinline__ u64 fastf_hash_mix(u64 h)
{
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
    return h;
}
template <class Acc> __device__ __forceinline__ u64 fastf_hash_bytes(const Acc &A, u64 off, u32 len, u64 h)
{
    for (u32 i = 0; i < len; i++) { h ^= A.byte(off + i); h *= 1099511628211ull; }
    return fastf_hash_mix(h ^ ((u64)len << 56));
}

// returns 0 = tag A absent, 1 = entry produced, 2 = corrupt record (stop the block), 3 = tag present but of the wrong type
template <class Acc>
__device__ __forceinline__ u32 fastf_tag_record(const Acc &A, u64 rec, u64 rend, u32 bs, const FastfTagQuery &Q, u64 *key, u64 *loc_a, u64 *loc_b)
{
    const u32 l_read_name = A.byte(rec + 8);
    const u32 n_cigar = fastf_acc_u16(A, rec + 12);
    const i32 l_seq = (i32)fastf_acc_u32(A, rec + 16);
    const i64 aoff = 32 + (i64)l_read_name + 4 * (i64)n_cigar + (((i64)l_seq + 1) >> 1) + (i64)l_seq;
    if (l_seq < 0 || aoff > (i64)bs) return 2;
    FastfAuxHit ha = {0, 0, 0}, hb = {0, 0, 0};
    const u32 want = Q.b0 ? 3u : 1u;
    u32 found = 0;
    u64 q = rec + (u64)aoff;
    while (rend - q >= 3 && found != want) {
        u32 t0, t1, ty;
        if (rend - q >= 4) { const u32 hd = fastf_acc_4bytes(A, q); t0 = hd & 255u; t1 = (hd >> 8) & 255u; ty = (hd >> 16) & 255u; }
        else { t0 = A.byte(q); t1 = A.byte(q + 1); ty = A.byte(q + 2); }
        const u64 v = q + 3;
        u64 next;
        u32 vlen = 0;
        if (ty == 'Z' || ty == 'H') {
            const u64 s = fastf_find_nul(A, v, rend);
            if (s >= rend) break;   // unterminated: this and every later tag is invisible (htslib)
            vlen = (u32)(s - v);
            next = s + 1;
        } else {
            u64 sz;
            switch (ty) {
            case 'A': case 'c': case 'C': sz = 1; break;
            case 's': case 'S': sz = 2; break;
            case 'i': case 'I': case 'f': sz = 4; break;
            case 'd': sz = 8; break;
            case 'B': {
                if (rend - v < 5) { sz = ~0ull; break; }
                u32 sub = A.byte(v);
                u64 cnt = fastf_acc_u32(A, v + 1);
                u64 es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : (sub == 'i' || sub == 'I' || sub == 'f') ? 4 : 0;
                sz = es ? 5 + es * cnt : ~0ull;
                break;
            }
            default: sz = ~0ull;
            }
            if (sz == ~0ull || sz > rend - v) break;
            next = v + sz;
        }
        if (t0 == Q.a0 && t1 == Q.a1 && !(found & 1u)) { ha.off = v; ha.len = vlen; ha.type = ty; found |= 1u; }
        else if (Q.b0 && t0 == Q.b0 && t1 == Q.b1 && !(found & 2u)) { hb.off = v; hb.len = vlen; hb.type = ty; found |= 2u; }
        q = next;
    }
    if (!(found & 1u)) return 0;
    if (Q.mode == FASTF_TAG_MODE_INT) {
        i64 x = 0;
        switch (ha.type) {
        case 'c': x = (i64)(int8_t)A.byte(ha.off); break;
        case 'C': x = A.byte(ha.off); break;
        case 's': x = (i64)(int16_t)fastf_acc_u16(A, ha.off); break;
        case 'S': x = fastf_acc_u16(A, ha.off); break;
        case 'i': x = (i64)(i32)fastf_acc_u32(A, ha.off); break;
        case 'I': x = fastf_acc_u32(A, ha.off); break;
        default: x = 0;
        }
        *key = (u64)(u32)x;   // printed with "%d": the low 32 bits, signed
        *loc_a = 0;
        *loc_b = 0;
        return 1;
    }
    // string mode: the reference hands bam_aux2Z() straight to strcmp/strcpy -- NULL for any other type, and for an absent CR
    if (!(ha.type == 'Z' || ha.type == 'H')) return 3;
    if (Q.b0 && (!(found & 2u) || !(hb.type == 'Z' || hb.type == 'H'))) return 3;
    if (ha.len > 0xffffu || hb.len > 0xffffu) return 3;
    u64 h = fastf_hash_bytes(A, ha.off, ha.len, 14695981039346656037ull ^ Q.seed);
    if (Q.b0) h = fastf_hash_bytes(A, hb.off, hb.len, h);
    *key = h & Q.key_mask;
    *loc_a = (ha.off << 16) | ha.len;
    *loc_b = Q.b0 ? ((hb.off << 16) | hb.len) : 0;
    return 1;
}

// stage layout: three planes of `stage_plane` u64 each (key, loc A, loc B), block b's slice at stage_off[b]
__global__ void __launch_bounds__(FASTF_PARSE_WARPS * 32)
fastf_bam_tags_kernel(const u8 *__restrict__ infl, u64 infl_total, const u64 *__restrict__ blk_off, const u32 *__restrict__ blk_isize, u32 nblocks, const u64 *__restrict__ first_record_off_ptr,
                      FastfTagQuery Q, const u64 *__restrict__ stage_off, u64 *__restrict__ stage, u64 stage_plane, u32 *__restrict__ blk_nrec, u32 *__restrict__ blk_nhit,
                      u32 *__restrict__ blk_status)
{
    __shared__ __align__(16) u8 s_win[FASTF_PARSE_WARPS][FASTF_PARSE_WIN];
    __shared__ __align__(8) u64 s_mbar[FASTF_PARSE_WARPS];   // one TMA completion barrier per warp (its window is private)
    const u32 lane = threadIdx.x & 31u;
    const u32 b = blockIdx.x * FASTF_PARSE_WARPS + (threadIdx.x >> 5);
    if (b >= nblocks) return;
    u8 *W = s_win[threadIdx.x >> 5];
    u64 *mbar = &s_mbar[threadIdx.x >> 5];
    u32 mbar_parity = 0;
    if (lane == 0) fastf_mbar_init(mbar, 1);
    __syncwarp();
    const u64 bstart = blk_off[b], bend = bstart + blk_isize[b];
    const u64 first_record_off = *first_record_off_ptr;
    u64 p = bstart > first_record_off ? bstart : first_record_off;
    u64 *out = stage + stage_off[b];
    const u64 cap = stage_off[b + 1] - stage_off[b];   // stage_off has nblocks + 1 entries; only a wrong record-start guess (FASTF_BAM_STRADDLE) can exceed it
    u32 nrec = 0, nhit = 0, status = 0;

    while (p < bend) {
        if (bend - p < 4) { status |= FASTF_ST_REC_STRADDLE; break; }
        const u64 wbase = p & ~15ull;
        u64 wend = wbase + FASTF_PARSE_WIN < infl_total ? wbase + FASTF_PARSE_WIN : infl_total;
        {
            const u64 need = ((bend + 15ull) & ~15ull) < wend ? ((bend + 15ull) & ~15ull) : wend;
            // the window arrives as ONE bulk copy of the TMA engine (a loop of 16-byte loads kept the warp waiting on DRAM round trips:
            // 28 % of the kernel's stall samples in round 1)
            __syncwarp();
            if (lane == 0) fastf_tma_load_1d(W, infl + wbase, (u32)(need - wbase), mbar);
            fastf_mbar_wait(mbar, mbar_parity);
            mbar_parity ^= 1u;
            __syncwarp();
            wend = need;
        }
        const FastfWinAcc WA = {W, wbase};
        u64 myrec = 0;
        u32 mybs = 0, nb = 0, stop = 0;
        u64 q = p;
        while (nb < 32u && q < bend) {
            if (bend - q < 4) { stop = FASTF_ST_REC_STRADDLE; break; }
            if (q + 4 > wend) break;
            const u32 bs = fastf_acc_u32(WA, q);
            if ((i32)bs < 32) { stop = FASTF_ST_REC_CORRUPT; break; }
            if (q + 4 + (u64)bs > bend) { stop = FASTF_ST_REC_STRADDLE; break; }
            if (q + 4 + (u64)bs > wend) break;
            if (lane == nb) { myrec = q + 4; mybs = bs; }
            nb++;
            q += 4 + (u64)bs;
        }
        u64 key = 0, la = 0, lb = 0;
        u32 r = 0;
        if (nb == 0 && !stop) {
            // the record at p is larger than the window: lane 0 walks it in global memory
            const u32 bs = fastf_ld_u32(infl + p);
            if ((i32)bs < 32) { status |= FASTF_ST_REC_CORRUPT; break; }
            if (p + 4 + (u64)bs > bend) { status |= FASTF_ST_REC_STRADDLE; break; }
            const FastfGlobAcc GA = {infl};
            if (lane == 0) r = fastf_tag_record(GA, p + 4, p + 4 + bs, bs, Q, &key, &la, &lb);
            nb = 1;
            q = p + 4 + (u64)bs;
        } else if (lane < nb) {
            r = fastf_tag_record(WA, myrec, myrec + mybs, mybs, Q, &key, &la, &lb);
        }
        const u32 badm = __ballot_sync(FASTF_FULL_MASK, r == 2);
        u32 good = nb;
        if (badm) good = (u32)__ffs((int)badm) - 1u;
        if (__ballot_sync(FASTF_FULL_MASK, r == 3 && lane < good)) status |= FASTF_ST_TAG_TYPE;
        const u32 cm = __ballot_sync(FASTF_FULL_MASK, r == 1 && lane < good);
        if (nhit + (u32)__popc(cm) > cap) { status |= FASTF_ST_REC_CORRUPT; break; }
        if (r == 1 && lane < good) {
            const u32 at = nhit + (u32)__popc(cm & fastf_lanemask_lt());
            out[at] = key;
            out[stage_plane + at] = la;
            out[2 * stage_plane + at] = lb;
        }
        nhit += (u32)__popc(cm);
        nrec += good;
        if (badm) { status |= FASTF_ST_REC_CORRUPT; break; }
        p = q;
        if (stop) { status |= stop; break; }
    }
    if (lane == 0) { blk_nrec[b] = nrec; blk_nhit[b] = nhit; blk_status[b] = status; }
}

// sorted[j] / perm[j] = key and ordinal of the j-th entry after the stable sort.  Neighbours with equal keys must hold equal
// bytes (transitively: equal to the group's first entry); *collision is set otherwise.
__global__ void __launch_bounds__(256)
fastf_taghist_verify_kernel(const u8 *__restrict__ infl, const u64 *__restrict__ sorted, const u32 *__restrict__ perm, const u64 *__restrict__ loc_a, const u64 *__restrict__ loc_b, u64 n,
                            u32 *__restrict__ collision)
{
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0 || j >= n || sorted[j] != sorted[j - 1]) return;
    const u32 x = perm[j], y = perm[j - 1];
    bool same = true;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const u64 lx = k ? loc_b[x] : loc_a[x], ly = k ? loc_b[y] : loc_a[y];
        if ((lx & 0xffffu) != (ly & 0xffffu)) { same = false; break; }
        const u8 *px = infl + (lx >> 16), *py = infl + (ly >> 16);
        for (u32 i = 0; i < (u32)(lx & 0xffffu); i++) if (px[i] != py[i]) { same = false; break; }
    }
    if (!same) *collision = 1u;
}

// one warp per group: representative (= first occurrence) value bytes -> blob; rep_* = loc of the group's first entry
__global__ void __launch_bounds__(256)
fastf_taghist_reps_kernel(const u32 *__restrict__ grp_val, const u64 *__restrict__ loc_a, const u64 *__restrict__ loc_b, u32 ngroups, u64 *__restrict__ rep_a, u64 *__restrict__ rep_b)
{
    const u32 g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    rep_a[g] = loc_a[grp_val[g]];
    rep_b[g] = loc_b[grp_val[g]];
}
__global__ void __launch_bounds__(256)
fastf_taghist_strings_kernel(const u8 *__restrict__ infl, const u64 *__restrict__ rep_a, const u64 *__restrict__ rep_b, const u64 *__restrict__ blob_off, u32 ngroups, u8 *__restrict__ blob)
{
    const u32 g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (g >= ngroups) return;
    const u64 la = rep_a[g], lb = rep_b[g];
    const u32 na = (u32)(la & 0xffffu), nb = (u32)(lb & 0xffffu);
    u8 *dst = blob + blob_off[g];
    for (u32 i = lane; i < na; i += 32) dst[i] = infl[(la >> 16) + i];
    for (u32 i = lane; i < nb; i += 32) dst[na + i] = infl[(lb >> 16) + i];
}
