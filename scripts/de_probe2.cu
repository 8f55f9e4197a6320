// probe: B200 hardware decompression engine (cuMemBatchDecompressAsync, DEFLATE) on the raw-deflate payloads of a BGZF file.
// usage: de_probe2 file.bam [max_blocks]     -- verifies every block against zlib and reports throughput
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <zlib.h>
#include "../fastf_b200/csrc/bgzf_index.h"
#define CKD(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char *s_; cuGetErrorString(r_, &s_); printf("FAIL %s -> %d %s\n", #x, (int)r_, s_ ? s_ : "?"); return 1; } } while (0)
#define CKR(x) do { cudaError_t r_ = (x); if (r_ != cudaSuccess) { printf("FAIL %s -> %s\n", #x, cudaGetErrorString(r_)); return 1; } } while (0)
int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    fseek(f, 0, SEEK_END); size_t n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> file(n);
    if (fread(file.data(), 1, n, f) != n) return 2;
    fclose(f);
    std::vector<FastfBgzfBlock> blocks; size_t used = 0;
    int rc = fastf_bgzf_index(file.data(), n, 0, blocks, &used);
    printf("index rc=%d blocks=%zu\n", rc, blocks.size());
    size_t nb = blocks.size();
    if (argc > 2 && (size_t)atol(argv[2]) < nb) nb = atol(argv[2]);
    CKR(cudaSetDevice(0));
    CKR(cudaFree(0));
    uint8_t *dcomp, *dout; uint32_t *dact;
    std::vector<uint64_t> out_off(nb); uint64_t total = 0;
    for (size_t i = 0; i < nb; i++) { out_off[i] = total; total += blocks[i].isize; total = (total + 15) & ~15ull; }
    CKR(cudaMalloc(&dcomp, n + 64)); CKR(cudaMalloc(&dout, total + 64)); CKR(cudaMalloc(&dact, nb * 4 + 4));
    CKR(cudaMemcpy(dcomp, file.data(), n, cudaMemcpyHostToDevice));
    int cap = -1;
    cuPointerGetAttribute(&cap, CU_POINTER_ATTRIBUTE_IS_HW_DECOMPRESS_CAPABLE, (CUdeviceptr)dcomp);
    printf("cudaMalloc pointer hw-decompress capable: %d\n", cap);
    std::vector<CUmemDecompressParams> P(nb);
    memset(P.data(), 0, nb * sizeof(CUmemDecompressParams));
    for (size_t i = 0; i < nb; i++) {
        P[i].srcNumBytes = blocks[i].in_len; P[i].dstNumBytes = blocks[i].isize; P[i].dstActBytes = dact + i;
        P[i].src = dcomp + blocks[i].in_off; P[i].dst = dout + out_off[i]; P[i].algo = CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE;
    }
    cudaStream_t s; CKR(cudaStreamCreate(&s));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    // 1. a single small batch first
    size_t erri = 0;
    size_t first = nb < 4 ? nb : 4;
    CKD(cuMemBatchDecompressAsync(P.data(), first, 0, &erri, s));
    CKR(cudaStreamSynchronize(s));
    printf("first batch of %zu ok\n", first);
    // 2. everything, timed, 3 times
    for (int rep = 0; rep < 3; rep++) {
        CKR(cudaMemsetAsync(dout, 0xAA, total, s));
        cudaEventRecord(e0, s);
        CKD(cuMemBatchDecompressAsync(P.data(), nb, 0, &erri, s));
        cudaEventRecord(e1, s);
        CKR(cudaStreamSynchronize(s));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        uint64_t inb = 0; for (size_t i = 0; i < nb; i++) inb += blocks[i].in_len;
        printf("rep %d: %zu blocks, %.1f MB in, %.1f MB out in %.3f ms -> %.1f GB/s out, %.1f GB/s in+out\n", rep, nb, inb / 1e6, total / 1e6, ms, total / ms / 1e6, (total + inb) / ms / 1e6);
    }
    // 3. verify
    std::vector<uint8_t> out(total + 64); std::vector<uint32_t> act(nb);
    CKR(cudaMemcpy(out.data(), dout, total, cudaMemcpyDeviceToHost)); CKR(cudaMemcpy(act.data(), dact, nb * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    std::vector<uint8_t> ref(65536 + 16);
    for (size_t i = 0; i < nb; i++) {
        z_stream zs; memset(&zs, 0, sizeof zs); inflateInit2(&zs, -15);
        zs.next_in = file.data() + blocks[i].in_off; zs.avail_in = blocks[i].in_len; zs.next_out = ref.data(); zs.avail_out = 65536;
        inflate(&zs, Z_FINISH); inflateEnd(&zs);
        if (act[i] != blocks[i].isize || memcmp(ref.data(), out.data() + out_off[i], blocks[i].isize)) { if (bad < 5) printf("block %zu mismatch act=%u isize=%u src_align=%zu\n", i, act[i], blocks[i].isize, (size_t)(blocks[i].in_off & 15)); bad++; }
    }
    printf("verify: %zu of %zu blocks differ\n", bad, nb);
    return bad ? 1 : 0;
}
