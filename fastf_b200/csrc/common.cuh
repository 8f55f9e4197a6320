// Common device/host helpers for the fastF B200 kernels (sm_100a only).
//
// Kernels are written in plain CUDA C++.  The only indirection is FASTF_LAUNCH (the <<<>>> launch
// syntax) and a handful of inline wrappers, so that tests/emu/cuda_emu.h can compile the very same
// kernel sources for the host and run them under a cooperative SIMT emulator (this container has no
// GPU; the emulator is test infrastructure and is never part of the product build).
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#ifdef FASTF_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#define FASTF_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;
typedef int64_t i64;

#define FASTF_FULL_MASK 0xffffffffu
#define FASTF_INVALID_KEY 0xffffffffffffffffull

// ---- per-block status / error bits (shared by inflate and parse kernels, surfaced through the C-ABI) ----
enum : u32 {
    FASTF_ST_OK = 0,
    FASTF_ST_BAD_BTYPE = 1u << 0,        // reserved deflate block type 3
    FASTF_ST_BAD_STORED = 1u << 1,       // stored block LEN != ~NLEN or overruns
    FASTF_ST_BAD_CODELENS = 1u << 2,     // over-subscribed / malformed dynamic header
    FASTF_ST_BAD_SYMBOL = 1u << 3,       // undecodable Huffman code / invalid length or distance symbol
    FASTF_ST_BAD_DISTANCE = 1u << 4,     // distance reaches before the start of the block
    FASTF_ST_OUT_OVERFLOW = 1u << 5,     // output would exceed ISIZE
    FASTF_ST_SIZE_MISMATCH = 1u << 6,    // decoded length != ISIZE
    FASTF_ST_IN_OVERRUN = 1u << 7,       // consumed more compressed bytes than the block holds
    FASTF_ST_REC_STRADDLE = 1u << 8,     // BAM record crosses the end of its BGZF block (foreign writer)
    FASTF_ST_REC_CORRUPT = 1u << 9,      // block_size < 32 or aux offset beyond the record
    FASTF_ST_UMI_TOO_LONG = 1u << 10,    // UB longer than the key layout allows
    FASTF_ST_AUX_CORRUPT = 1u << 11,     // malformed aux field (htslib: treated as "tag absent")
    FASTF_ST_BAD_HEADER = 1u << 12,      // BAM magic / header does not fit the first chunk
    FASTF_ST_BAD_CRC = 1u << 13,         // CRC-32 of the inflated block differs from the BGZF trailer
    FASTF_ST_TAG_TYPE = 1u << 14,        // crb / extract: a tag the reference reads with bam_aux2Z() is not a string (or CR is absent): it dereferences NULL there
};

// dynamic shared memory of the running CTA
#ifdef FASTF_EMU
#define FASTF_DYN_SMEM(ptr) u8 *ptr = emu::dyn_smem()
#else
#define FASTF_DYN_SMEM(ptr) extern __shared__ __align__(16) u8 fastf_dyn_smem_[]; u8 *ptr = fastf_dyn_smem_
#endif

// polite spin while another warp of the CTA makes progress.  fastf_spin_pause: the WHOLE warp has nothing to do (sleeps);
// fastf_spin_poll: only some lanes of a warp wait while others work -- a sleep there would stall the working lanes at the next
// reconvergence point, so on the GPU it is a no-op (the emulator still has to let the other fibers run).
#ifdef FASTF_EMU
__device__ __forceinline__ void fastf_spin_pause() { emu::spin_yield(); }
__device__ __forceinline__ void fastf_spin_poll() { emu::spin_yield(); }
#else
#ifndef FASTF_PAUSE_NS
#define FASTF_PAUSE_NS 400
#endif
__device__ __forceinline__ void fastf_spin_pause() { __nanosleep(FASTF_PAUSE_NS); }
__device__ __forceinline__ void fastf_spin_poll() {}
#endif

// ---- TMA 1-D bulk copy global -> shared memory, completion on an mbarrier (sm_90+/sm_100a: SASS UBLKCP + SYNCS) ----
// One lane issues the copy of a whole window; the TMA engine streams it without occupying the warp's load/store slots and every
// lane then waits on the barrier's phase.  dst, src and bytes are multiples of 16.  The emulator build copies synchronously.
#ifdef FASTF_EMU
__device__ __forceinline__ void fastf_mbar_init(u64 *, u32) {}
__device__ __forceinline__ void fastf_tma_load_1d(void *dst_smem, const void *src_gmem, u32 bytes, u64 *) { memcpy(dst_smem, src_gmem, bytes); }
__device__ __forceinline__ void fastf_mbar_wait(u64 *, u32) {}
#else
__device__ __forceinline__ u32 fastf_smem_addr(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fastf_mbar_init(u64 *mbar, u32 arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fastf_smem_addr(mbar)), "r"(arrivals));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// the calling thread is the barrier's one arrival: it announces the byte count, then starts the copy
__device__ __forceinline__ void fastf_tma_load_1d(void *dst_smem, const void *src_gmem, u32 bytes, u64 *mbar)
{
    const u32 d = fastf_smem_addr(dst_smem), b = fastf_smem_addr(mbar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic-proxy accesses of the window are ordered before the engine's writes
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src_gmem), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void fastf_mbar_wait(u64 *mbar, u32 parity)
{
    const u32 b = fastf_smem_addr(mbar);
    u32 done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b), "r"(parity) : "memory");
    }
}
#endif

__device__ __forceinline__ u32 fastf_lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ u32 fastf_lanemask_lt() { return (1u << (threadIdx.x & 31u)) - 1u; }

// unaligned little-endian reads from global memory
__device__ __forceinline__ u32 fastf_ld_u16(const u8 *p) { return (u32)p[0] | ((u32)p[1] << 8); }
__device__ __forceinline__ u32 fastf_ld_u32(const u8 *p) { return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24); }

// ---- block-wide exclusive scan (u32) for THREADS = multiple of 32, <= 1024 ----
// returns the exclusive prefix of v over the CTA in thread order; *total gets the CTA sum.
template <int THREADS>
__device__ __forceinline__ u32 fastf_block_exscan(u32 v, u32 *total)
{
    __shared__ u32 s_warp[THREADS / 32];
    __shared__ u32 s_total;
    const u32 lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    u32 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(FASTF_FULL_MASK, inc, o);
        if ((int)lane >= o) inc += t;
    }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        u32 x = (lane < THREADS / 32) ? s_warp[lane] : 0u;
        u32 xi = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(FASTF_FULL_MASK, xi, o);
            if ((int)lane >= o) xi += t;
        }
        if (lane < THREADS / 32) s_warp[lane] = xi - x;
        if (lane == 31) s_total = xi;
    }
    __syncthreads();
    u32 r = s_warp[w] + inc - v;
    *total = s_total;
    __syncthreads();   // s_warp / s_total may be reused by the next call
    return r;
}
