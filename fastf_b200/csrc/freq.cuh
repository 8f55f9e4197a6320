// K5 -- `freq`: barcode(+UMI) histogram of an R1 FASTQ on device.
//
// Replaces cell_counts() (reference src/count.c:3-21): per record, key = first l+u bytes of the
// sequence line (get_fastq src/filter.c:15-37, substring src/filter.c:260-275) inserted into a BST
// histogram (insert_tree src/filter.c:105-124).  On device: newline scan over the inflated text
// (records are lines 4r..4r+3, so the sequence line of record r starts after newline 4r), 2-bit pack
// of pure-ACGT keys (numeric order == strcmp order because A<C<G<T), stable radix sort of
// (key, read ordinal), run-length encode -> (key, count, first-occurrence ordinal).  The reference's
// output order (BST pre-order, src/filter.c:139-148) is a function of exactly those three columns.
// Reads whose key holds anything but ACGT (N, a short line's '\n', ...) are exported raw to the host,
// which merges them by byte order -- 2-bit packing is not order preserving for them (A<C<G<N<T).
#pragma once
#include "common.cuh"

#define FASTF_NL_THREADS 256
#define FASTF_NL_BYTES_PER_THREAD 16
#define FASTF_NL_TILE (FASTF_NL_THREADS * FASTF_NL_BYTES_PER_THREAD)
#define FASTF_FREQ_SENTINEL 0xffffffffffffffffull
#define FASTF_FREQ_MAX_KEY 31        // bases; 2 bits each, one spare bit pattern for the sentinel
#define FASTF_FREQ_EXC_STRIDE 32     // bytes kept per exceptional read

// newlines among the 16 bytes at off (as a bit mask); bytes in front of `skip` do not exist (the text of a streamed chunk starts at
// text + skip so that text itself stays 16-byte aligned)
__device__ __forceinline__ u32 fastf_count_nl16(const u8 *__restrict__ text, u64 off, u64 n, u32 *mask, u64 skip = 0)
{
    u32 m = 0;
    if (off + 16 <= n && ((off & 15u) == 0)) {
        const uint4 v = *reinterpret_cast<const uint4 *>(text + off);
        const u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
            for (int b = 0; b < 4; b++)
                if (((w[k] >> (8 * b)) & 0xffu) == '\n') m |= 1u << (4 * k + b);
    } else {
        for (u32 k = 0; k < 16; k++) if (off + k < n && text[off + k] == '\n') m |= 1u << k;
    }
    if (off < skip) m = (skip - off >= 16u) ? 0u : (m & ~((1u << (u32)(skip - off)) - 1u));
    *mask = m;
    return (u32)__popc(m);
}

__global__ void __launch_bounds__(FASTF_NL_THREADS) fastf_nl_count_kernel(const u8 *__restrict__ text, u64 n, u32 *__restrict__ tile_counts, u64 skip)
{
    __shared__ u32 s_c;
    if (threadIdx.x == 0) s_c = 0;
    __syncthreads();
    const u64 off = (u64)blockIdx.x * FASTF_NL_TILE + (u64)threadIdx.x * FASTF_NL_BYTES_PER_THREAD;
    u32 m;
    u32 c = off < n ? fastf_count_nl16(text, off, n, &m, skip) : 0u;
    c = __reduce_add_sync(FASTF_FULL_MASK, c);
    if ((threadIdx.x & 31u) == 0) atomicAdd(&s_c, c);
    __syncthreads();
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = s_c;
}

// For every newline with global index j = 4r: pack the key of record r.
// exc_count counts ALL exceptional reads; only the first exc_cap are stored (host re-runs with a larger cap on overflow).
// Streaming (the text is one chunk of a larger file, preceded by the tail of the previous chunk): nl_base = global index of the first
// newline at or behind `skip`; a key whose FASTF_FREQ_EXC_STRIDE bytes end at or before `done_below` was emitted by the previous
// chunk, one whose bytes run past n waits for the next chunk unless this is the last one.
__global__ void __launch_bounds__(FASTF_NL_THREADS)
fastf_freq_keys_kernel(const u8 *__restrict__ text, u64 n, const u32 *__restrict__ tile_off, u32 klen, u64 *__restrict__ keys, u64 n_keys_cap,
                       u32 *__restrict__ exc_count, u32 exc_cap, u32 *__restrict__ exc_ord, u8 *__restrict__ exc_bytes, u64 skip, u64 nl_base, u64 done_below, u32 last_chunk)
{
    const u64 off = (u64)blockIdx.x * FASTF_NL_TILE + (u64)threadIdx.x * FASTF_NL_BYTES_PER_THREAD;
    u32 m = 0;
    u32 c = off < n ? fastf_count_nl16(text, off, n, &m, skip) : 0u;
    u32 tot;
    u64 j = nl_base + (u64)tile_off[blockIdx.x] + fastf_block_exscan<FASTF_NL_THREADS>(c, &tot);
    while (m) {
        const u32 k = (u32)__ffs((int)m) - 1u;
        m &= m - 1u;
        const u64 start = off + k + 1;
        if ((j & 3ull) == 0 && start + FASTF_FREQ_EXC_STRIDE > done_below && (last_chunk || start + FASTF_FREQ_EXC_STRIDE <= n)) {
            const u64 r = j >> 2;
            u64 key = 0;
            bool good = true;
            for (u32 cidx = 0; cidx < klen; cidx++) {
                const u64 a = start + cidx;
                const u32 ch = a < n ? text[a] : 0u;
                const u32 code = ch == 'A' ? 0u : ch == 'C' ? 1u : ch == 'G' ? 2u : ch == 'T' ? 3u : 4u;
                if (code > 3u) { good = false; break; }
                key = (key << 2) | code;
            }
            if (r < n_keys_cap) {
                if (good) {
                    keys[r] = key;
                } else {
                    keys[r] = FASTF_FREQ_SENTINEL;
                    const u32 e = atomicAdd(exc_count, 1u);
                    if (e < exc_cap) {
                        exc_ord[e] = (u32)r;
                        for (u32 cidx = 0; cidx < FASTF_FREQ_EXC_STRIDE; cidx++) {
                            const u64 a = start + cidx;
                            exc_bytes[(u64)e * FASTF_FREQ_EXC_STRIDE + cidx] = a < n ? text[a] : 0;
                        }
                    }
                }
            }
        }
        j++;
    }
}

// ---- generic two-phase compaction of (key != sentinel) with the original index as payload ----
#define FASTF_CP_THREADS 256
#define FASTF_CP_ITEMS 8
#define FASTF_CP_TILE (FASTF_CP_THREADS * FASTF_CP_ITEMS)
__global__ void __launch_bounds__(FASTF_CP_THREADS) fastf_compact_count_kernel(const u64 *__restrict__ keys, u64 n, u32 *__restrict__ tile_counts)
{
    __shared__ u32 s_c;
    if (threadIdx.x == 0) s_c = 0;
    __syncthreads();
    const u64 base = (u64)blockIdx.x * FASTF_CP_TILE;
    u32 c = 0;
#pragma unroll
    for (int k = 0; k < FASTF_CP_ITEMS; k++) {
        u64 i = base + (u64)k * FASTF_CP_THREADS + threadIdx.x;
        if (i < n) c += keys[i] != FASTF_FREQ_SENTINEL;
    }
    c = __reduce_add_sync(FASTF_FULL_MASK, c);
    if ((threadIdx.x & 31u) == 0) atomicAdd(&s_c, c);
    __syncthreads();
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = s_c;
}
__global__ void __launch_bounds__(FASTF_CP_THREADS)
fastf_compact_scatter_kernel(const u64 *__restrict__ keys, u64 n, const u32 *__restrict__ tile_off, u64 *__restrict__ out_keys, u32 *__restrict__ out_idx)
{
    const u64 base = (u64)blockIdx.x * FASTF_CP_TILE + (u64)threadIdx.x * FASTF_CP_ITEMS;
    u64 k8[FASTF_CP_ITEMS];
    u32 flags = 0, c = 0;
#pragma unroll
    for (int k = 0; k < FASTF_CP_ITEMS; k++) {
        u64 i = base + k;
        bool v = false;
        if (i < n) { k8[k] = keys[i]; v = k8[k] != FASTF_FREQ_SENTINEL; }
        flags |= (u32)v << k;
        c += v;
    }
    u32 tot;
    u64 o = (u64)tile_off[blockIdx.x] + fastf_block_exscan<FASTF_CP_THREADS>(c, &tot);
#pragma unroll
    for (int k = 0; k < FASTF_CP_ITEMS; k++)
        if (flags & (1u << k)) { out_keys[o] = k8[k]; out_idx[o] = (u32)(base + k); o++; }
}
__global__ void __launch_bounds__(256) fastf_iota_kernel(u32 *__restrict__ out, u64 n)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (u32)i;
}
