// K1c -- CRC-32 of every inflated BGZF block against the CRC32 field of its trailer.
//
// htslib checks this inside bgzf_read_block (the reference reaches it through sam_read1, src/bam2db_ds.c:360; the FASTQ side
// through zlib's gzread, src/filter.c:22-27): a block whose payload inflates cleanly but to the wrong bytes ends the read loop
// there.  Here such a block raises FASTF_ST_BAD_CRC and the job fails loudly instead of producing a truncated result.
//
// One warp per block (grid-stride).  CRC-32 (IEEE 802.3, reflected, polynomial 0xEDB88320) is linear over GF(2), so the 32
// lanes checksum 32 contiguous segments independently -- slice-by-4 table lookups out of shared memory (one private copy of
// the tables per lane, entry i of lane l at word 32 i + l, so that the 32 data-dependent lookups of a warp never collide on a
// bank), 16-byte loads -- and
// the lane results are folded by a 5-level tree: crc(A || B) = crc(A) * x^(8 |B|) mod P  xor  crc(B).  All segments but the
// first have the same length S, so level k needs one constant, x^(8 S 2^k) mod P, built from a per-CTA table of
// x^(8 2^j) mod P.  Bound: HBM read of the inflated bytes (once); measured in profiles/README.md.
#pragma once
#include "common.cuh"

#define FASTF_CRC_POLY 0xEDB88320u
#define FASTF_CRC_WARPS 32            // one CTA per SM: the lane-replicated tables take 128 KB of shared memory

// a * b mod P, polynomials in the reflected representation (bit 31 = x^0), as in zlib's multmodp
__device__ __forceinline__ u32 fastf_crc_mulmod(u32 a, u32 b)
{
    u32 p = 0;
#pragma unroll 8
    for (int i = 0; i < 32; i++) {
        if (a & (0x80000000u >> i)) p ^= b;
        b = (b >> 1) ^ ((b & 1u) ? FASTF_CRC_POLY : 0u);
    }
    return p;
}

struct FastfCrcTables {
    u32 t[4][256][32];   // slice-by-4, replicated per lane
    u32 x8pow[24];       // x^(8 * 2^j) mod P
};

__device__ __forceinline__ void fastf_crc_tables_init(FastfCrcTables &T)
{
    for (u32 i = threadIdx.x; i < 256; i += blockDim.x) {
        u32 c = i;
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ ((c & 1u) ? FASTF_CRC_POLY : 0u);
        T.t[0][i][0] = c;
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < 256; i += blockDim.x) {
        u32 c = T.t[0][i][0];
        for (int k = 1; k < 4; k++) {
            c = T.t[0][c & 0xffu][0] ^ (c >> 8);
            T.t[k][i][0] = c;
        }
    }
    if (threadIdx.x == 0) {
        u32 p = 0x00800000u;   // x^8
        for (int j = 0; j < 24; j++) { T.x8pow[j] = p; p = fastf_crc_mulmod(p, p); }
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < 4u * 256u * 32u; i += blockDim.x)
        if (i & 31u) (&T.t[0][0][0])[i] = (&T.t[0][0][0])[i & ~31u];
    __syncthreads();
}

__device__ __forceinline__ u32 fastf_crc_byte(const FastfCrcTables &T, u32 lane, u32 c, u32 b) { return T.t[0][(c ^ b) & 0xffu][lane] ^ (c >> 8); }
__device__ __forceinline__ u32 fastf_crc_word(const FastfCrcTables &T, u32 lane, u32 c, u32 w)
{
    c ^= w;
    return T.t[3][c & 0xffu][lane] ^ T.t[2][(c >> 8) & 0xffu][lane] ^ T.t[1][(c >> 16) & 0xffu][lane] ^ T.t[0][c >> 24][lane];
}

// register state after running `n` bytes at p through the CRC starting from state c (no pre/post inversion)
__device__ __forceinline__ u32 fastf_crc_run(const FastfCrcTables &T, u32 lane, u32 c, const u8 *p, u32 n)
{
    while (n && ((uintptr_t)p & 15u)) { c = fastf_crc_byte(T, lane, c, *p++); n--; }
    for (; n >= 16; n -= 16, p += 16) {
        const uint4 v = *reinterpret_cast<const uint4 *>(p);
        c = fastf_crc_word(T, lane, c, v.x);
        c = fastf_crc_word(T, lane, c, v.y);
        c = fastf_crc_word(T, lane, c, v.z);
        c = fastf_crc_word(T, lane, c, v.w);
    }
    while (n) { c = fastf_crc_byte(T, lane, c, *p++); n--; }
    return c;
}

// comp/in_off/in_len locate the block trailers (CRC32 sits right behind the payload); infl/out_off/isize the inflated bytes.
// Blocks that already carry an inflate error are skipped (their bytes are not meaningful).
__global__ void __launch_bounds__(FASTF_CRC_WARPS * 32) fastf_bgzf_crc32_kernel(const u8 *comp, u64 comp_total, const u64 *in_off, const u32 *in_len, const u8 *infl, const u64 *out_off,
                                                                              const u32 *isize, u32 nblocks, u32 *status)
{
    FASTF_DYN_SMEM(smem);
    FastfCrcTables &T = *reinterpret_cast<FastfCrcTables *>(smem);
    fastf_crc_tables_init(T);
    const u32 lane = threadIdx.x & 31u;
    const u32 warp0 = blockIdx.x * FASTF_CRC_WARPS + (threadIdx.x >> 5), nwarps = gridDim.x * FASTF_CRC_WARPS;
    for (u32 b = warp0; b < nblocks; b += nwarps) {
        if (status[b] != 0) continue;
        const u32 n = isize[b];
        const u8 *src = infl + out_off[b];
        // lane 0: [0, n - 31 S); lane i >= 1: the i-th of the 31 segments of S bytes that end the block
        const u32 S = (n >> 5) & ~15u;
        const u32 first = n - 31u * S;
        const u32 beg = lane ? first + (lane - 1u) * S : 0u, len = lane ? S : first;
        u32 c = fastf_crc_run(T, lane, lane ? 0u : 0xffffffffu, src + beg, len);
        if (S) {
            // lane k < 5 builds x^(8 S 2^k): product of x8pow[j + k] over the set bits j of S
            u32 pw = 0x80000000u;   // x^0
            for (u32 j = 4; j < 12; j++)
                if ((S >> j) & 1u) pw = fastf_crc_mulmod(pw, T.x8pow[j + (lane < 5 ? lane : 0u)]);
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const u32 shiftk = __shfl_sync(FASTF_FULL_MASK, pw, k);
                const u32 right = __shfl_down_sync(FASTF_FULL_MASK, c, 1u << k);
                // lanes whose index is a multiple of 2^(k+1) absorb the neighbour 2^k to the right (|right| = S 2^k bytes)
                if ((lane & ((2u << k) - 1u)) == 0) c = fastf_crc_mulmod(c, shiftk) ^ right;
            }
        }
        if (lane == 0) {
            const u64 t = in_off[b] + in_len[b];
            u32 want = 0, bad = 0;
            if (t + 4 > comp_total) bad = 1;
            else want = (u32)comp[t] | ((u32)comp[t + 1] << 8) | ((u32)comp[t + 2] << 16) | ((u32)comp[t + 3] << 24);
            if (bad || (c ^ 0xffffffffu) != want) status[b] = FASTF_ST_BAD_CRC;
        }
    }
}
