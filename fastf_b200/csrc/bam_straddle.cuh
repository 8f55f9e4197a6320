// K2s -- BAM files whose records straddle BGZF blocks (writers other than htslib: htsjdk, STAR's own BGZF layer).
//
// The reference never sees block boundaries (htslib's bgzf_read hands sam_read1 a byte stream, reference src/bam2db_ds.c:360).
// Here the blocks of a chunk are inflated back to back into one contiguous buffer, so a record may simply run on into the next
// block; what the per-block kernels (K2 parse, K6 tags) need is a RECORD START to begin at.  Opt-in mode FASTF_BAM_STRADDLE:
//   1. fastf_bam_guess_kernel: per BGZF block, the first offset in [block start, block end) from which three consecutive
//      plausible record headers follow (offset 0 of the block for htslib files); NONE for a block that lies inside one long record;
//   2. fastf_bam_virtual_blocks_kernel: virtual block k = [v_k, v_next), empty where there is no start;
//   3. K2 / K6 run unchanged over the virtual blocks and VERIFY the guesses: a chain of records must end exactly on its virtual
//      block's end (= the next guess), otherwise the block reports FASTF_ST_REC_STRADDLE / REC_CORRUPT and the job fails loudly.
//      By induction from the first record (end of the BAM header, known exactly) every guess that passes is a true record start.
// The mode keeps the whole file in one chunk (no record is cut by a chunk boundary): the inflated bytes must fit HBM.
#pragma once
#include "common.cuh"

#define FASTF_NO_START 0xffffffffffffffffull

// fixed part of a record at o: block_size | refID pos l_read_name mapq bin n_cigar flag l_seq next_refID next_pos tlen | read_name ...
__device__ __forceinline__ bool fastf_bam_header_plausible(const u8 *__restrict__ infl, u64 o, u64 n, u64 *next)
{
    if (o + 36 > n) return false;
    const u32 bs = fastf_ld_u32(infl + o);
    if (bs < 32u || bs > (1u << 27)) return false;
    const i32 ref_id = (i32)fastf_ld_u32(infl + o + 4), pos = (i32)fastf_ld_u32(infl + o + 8);
    const u32 l_read_name = infl[o + 12], n_cigar = fastf_ld_u16(infl + o + 16);
    const i32 l_seq = (i32)fastf_ld_u32(infl + o + 20);
    const i32 next_ref = (i32)fastf_ld_u32(infl + o + 24), next_pos = (i32)fastf_ld_u32(infl + o + 28);
    if (ref_id < -1 || ref_id >= (1 << 24) || next_ref < -1 || next_ref >= (1 << 24) || pos < -1 || next_pos < -1) return false;
    if (l_read_name == 0 || l_seq < 0 || l_seq > (1 << 26)) return false;
    const u64 fixed = 32ull + l_read_name + 4ull * n_cigar + (((u64)l_seq + 1) >> 1) + (u64)l_seq;
    if (fixed > bs) return false;
    if (o + 36 + l_read_name > n) return false;
    if (infl[o + 36 + l_read_name - 1] != 0) return false;                       // the read name is NUL-terminated ...
    if (l_read_name > 1 && infl[o + 36 + l_read_name - 2] == 0) return false;   // ... and holds no NUL before that
    *next = o + 4 + (u64)bs;
    return *next <= n;
}

// one warp per BGZF block: guess[k] = first plausible record start inside the block, FASTF_NO_START if none
__global__ void __launch_bounds__(256)
fastf_bam_guess_kernel(const u8 *__restrict__ infl, u64 infl_bytes, const u64 *__restrict__ blk_off, const u32 *__restrict__ blk_isize, u32 nblocks, const u64 *__restrict__ first_record_off_ptr,
                       u64 *__restrict__ guess)
{
    const u32 b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (b >= nblocks) return;
    const u64 bstart = blk_off[b], bend = bstart + blk_isize[b];
    const u64 fro = *first_record_off_ptr;   // end of the BAM header: the one record start that is known
    u64 found = FASTF_NO_START;
    if (bend <= fro) {
        // header bytes only
    } else if (bstart <= fro) {
        found = fro < infl_bytes ? fro : FASTF_NO_START;
    } else {
        for (u64 base = bstart; base < bend && found == FASTF_NO_START; base += 32) {
            const u64 o = base + lane;
            u64 nx = 0;
            bool ok = o < bend && fastf_bam_header_plausible(infl, o, infl_bytes, &nx);
            // two more headers behind it (the end of the file ends the chain)
            for (int k = 0; k < 2 && ok && nx < infl_bytes; k++) ok = fastf_bam_header_plausible(infl, nx, infl_bytes, &nx);
            const u32 m = __ballot_sync(FASTF_FULL_MASK, ok);
            if (m) found = base + (u32)__ffs((int)m) - 1u;
        }
    }
    if (lane == 0) guess[b] = found;
}

// virtual block k = [first start at or after block k, first start at or after block k+1) -- empty when block k holds no start
__global__ void __launch_bounds__(256)
fastf_bam_virtual_blocks_kernel(const u64 *__restrict__ guess, u32 nblocks, u64 infl_bytes, u64 *__restrict__ virt_off, u32 *__restrict__ virt_size, u32 *__restrict__ status)
{
    const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nblocks) return;
    u32 j = k;
    while (j < nblocks && guess[j] == FASTF_NO_START) j++;
    const u64 lo = j < nblocks ? guess[j] : infl_bytes;
    u32 j2 = k + 1;
    while (j2 < nblocks && guess[j2] == FASTF_NO_START) j2++;
    const u64 hi = j2 < nblocks ? guess[j2] : infl_bytes;
    virt_off[k] = lo;
    if (hi < lo || hi - lo > 0xffffffffull) { virt_size[k] = 0; atomicOr(status, (u32)FASTF_ST_REC_STRADDLE); return; }
    virt_size[k] = guess[k] == FASTF_NO_START ? 0u : (u32)(hi - lo);
}
